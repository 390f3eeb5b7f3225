// C ABI (include/loco_asr.h): handle, checkpoint ingestion, batch geometry and the encoder schedule.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <list>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/loco_asr.h"
#include "common.cuh"
#include "internal.h"

using namespace loco;

namespace {

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
    int64_t numel() const {
        int64_t n = 1;
        for (int64_t s : shape) n *= s;
        return n;
    }
};

struct LayerW {
    bf16 *wqkv, *wo, *w1, *w2;
    float *bqkv, *bo, *b1, *b2, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
    // deferred LayerNorm (GemmEpilogue EPI_LN_*): weights with the preceding LayerNorm's gamma folded in, their column sums
    // c1[n] = sum_k W'[n, k] (of the bf16-rounded W', so the mean term cancels exactly) and c2[n] = sum_k beta[k] W[n, k] + b[n]
    bf16 *wqkv_ln = nullptr, *w1_ln = nullptr;       // wqkv_ln: layers >= 1 (final_layer_norm of the layer before)
    float *qkv_c1 = nullptr, *qkv_c2 = nullptr, *ffn_c1 = nullptr, *ffn_c2 = nullptr;
    float *bo_ln = nullptr, *b2_ln = nullptr;        // out_proj / FFN2 bias plus the beta of the LayerNorm their residual goes through
};

struct Buf {
    size_t off = 0;
    int64_t rows = 0, cols = 0;
    int dtype = LOCO_BF16;
};

struct Layout {
    int n_utts = 0;
    int64_t R6 = 0, total_frames = 0, total_samples = 0;
    int max_t0 = 0, max_t6 = 0, max_slot6 = 0, chunks = 1;
    std::vector<UttMeta> meta;
    std::vector<PcTile> pc_tiles;
    std::vector<int32_t> c0_tile_start;         // conv0: first 128-frame tile of each utterance (n_utts + 1 entries; conv0_tc.cu)
    size_t off_wfold = 0;                       // conv0: per-utterance folded weights, kConv0FoldBytes each
    std::vector<int32_t> pp_map;                // positional conv: timeline frame -> row (internal.h, posconv_pp.cu)
    int n_vtiles = 0;
    size_t off_stats1 = 0, off_stats2 = 0;      // deferred LayerNorm: row statistics of attn_res / ffn_res, [R6, 6, 2] fp32
    size_t off_partial = 0, off_scale = 0, off_shift = 0, off_rowframe = 0;
    std::map<std::string, Buf> bufs;
    size_t bytes = 0;                           // caller-owned workspace: stage buffers and per-call scratch only
};

std::string g_create_error;

}  // namespace

// Arena for the blocks of implicit plans: a device chunk and a pinned host chunk of the same size, carved in step.  cudaMalloc /
// cudaMallocHost cost milliseconds of host time and can wait for the device; per plan that was a 2-3 ms hole in front of every
// first-time batch (bench full pass: 44.2 ms per step against 41.5 ms of kernels).  A chunk holds ~80 plans of a 131 k-frame batch.
struct PlanChunk {
    uint8_t* dev = nullptr;
    uint8_t* host = nullptr;
    size_t cap = 0, used = 0;
    int refs = 0;
};
constexpr size_t kPlanChunkBytes = (size_t)64 << 20;

// A batch geometry made ready to launch (loco_plan_create): the layout, the attention work lists, and one small device block
// (owned by the plan) holding everything the kernels read about the batch -- utterance metadata and tile lists.  Encoding with
// a plan enqueues kernels and memset nodes only: no host-to-device copy, no host synchronisation, hence capturable.
struct loco_batch_plan {
    int kind = 0;                               // 0 speech (lengths = samples), 1 text (lengths = tokens)
    std::vector<int32_t> lengths;
    Layout L;
    std::vector<PcTile> at_tiles;               // one-item tcgen05 kernel: 128-query tiles of the utterances routed to it
    std::vector<PcTile> at_tiles64;             // two-pipeline tcgen05 kernel: 64-query tiles of the utterances routed to it
    std::vector<int32_t> at_utts;               // LOCO_DEBUG builds: utterances routed to the mma.sync cross-check kernel
    int at_ms_max_t6 = 0;
    uint8_t* dev = nullptr;                     // [meta | pc_tiles | at_tiles | at_utts | pp_map]
    size_t d_ppmap = 0, d_c0tiles = 0;
    // plans made on behalf of loco_encode upload their block asynchronously on the encoding stream (no device synchronisation in
    // the middle of a queue of encodes): pinned staging copy kept for the plan's life, event for encodes on other streams
    uint8_t* staging = nullptr;
    struct PlanChunk* chunk = nullptr;          // the arena chunk `dev` and `staging` were carved from (implicit plans); null: own cudaMalloc
    cudaEvent_t uploaded = nullptr;
    cudaStream_t upload_stream = nullptr;
    size_t d_meta = 0, d_pctiles = 0, d_attiles = 0, d_attiles64 = 0, d_atutts = 0, dev_bytes = 0;
    int device = 0;
    bool cached = false;                        // owned by the handle's plan cache (loco_encode), not by the caller
    uint64_t knobs = 0;                         // debug-knob state the work lists were built under
};

struct loco_handle {
    loco_config cfg;
    int device = 0;
    int num_sms = 148;
    std::string err;
    std::map<std::string, HostTensor> host;
    bool finalized = false;
    std::vector<void*> allocs;
    // device weights
    float *w0 = nullptr, *gn_w = nullptr, *gn_b = nullptr;
    bf16* conv_w[8] = {};
    float *pln_w = nullptr, *pln_b = nullptr, *proj_b = nullptr;
    bf16* proj_w = nullptr;
    bf16* pos_w = nullptr;      // [g][tap][out][in]          (mma.sync debug kernel)
    bf16* pos_w_tc = nullptr;   // [g][tap][in/8][out][in%8]  (one-phase tcgen05 debug kernel: per-tap UMMA B operand)
    bf16* pos_w_pp = nullptr;   // [g][in/8][3 zero taps | tap | 8 zero taps][out][in%8]  (polyphase tcgen05 kernel)
    float* pos_b = nullptr;
    float* sin_table = nullptr;
    int sin_rows = 0;
    float *eln_w = nullptr, *eln_b = nullptr;
    bf16* pe_k = nullptr;
    // text prenet (SpeechT5TextEncoderPrenet): present when the state dict carried prenet.embed_tokens.weight
    bool has_speech = false, has_text = false;
    float* txt_embed = nullptr;   // [vocab, 768] fp32
    int txt_vocab = 0;
    float txt_alpha = 1.f;
    float* txt_pe = nullptr;      // [txt_pe_rows, 768] fp32, SpeechT5ScaledPositionalEncoding table
    int txt_pe_rows = 0;
    std::vector<LayerW> layers;
    // classifier head fused into the last kernel (loco_set_head / loco_set_head_outputs)
    HeadArgs head;
    bool head_set = false;
    // debug
    int gemm_impl = 2;          // 2 = tcgen05 CTA pair [default], 0 = tcgen05 single CTA, 1 = SIMT reference
    int posconv_impl = 0;       // 0 = polyphase tcgen05 [default], 1 = mma.sync, 2 = one-phase tcgen05
    int conv0_impl = 0;         // 0 = tcgen05 [default], 1 = mma.sync
    int ln_impl = 0;            // 0 = LayerNorms of the transformer layers deferred into the GEMM epilogues [default], 1 = LayerNorm kernels
    int attn_impl = 0;          // 0 = tcgen05 [default, the only product kernel], 1 = mma.sync cross-check, -1 = by length (round-1 routing)
    bool attn_p2 = true;        // utterances shorter than attn_p2_max_frames go to the two-pipeline tcgen05 kernel (attention_p2.cu),
    int attn_p2_max_frames = 193;   // the others to the one-item-per-SM tcgen05 kernel (attention_tc.cu),
    bool attn_p2_tail = true;       // except a tail of at most 64 query rows, which becomes one two-pipeline item
    int attn_tc_min_frames = 193;   // utterances with at least this many frames use the tcgen05 attention kernel,
    int attn_tc_lo = 76, attn_tc_hi = 128;   // ... and so do utterances that fill most of one 128-query tile (see loco_encode)
    alignas(64) CUtensorMap pe_map;   // pe_k [320, 64] for the tcgen05 attention kernel
    int stop_after_layer = -1;
    // plans built on behalf of loco_encode / loco_encode_text, keyed by the batch's lengths (LRU): a set that is encoded
    // batch by batch, epoch after epoch, plans each batch once
    std::vector<PlanChunk*> plan_chunks;
    std::list<loco_batch_plan*> plan_lru;
    std::unordered_map<uint64_t, std::list<loco_batch_plan*>::iterator> plan_index;
    const loco_batch_plan* last_plan = nullptr;       // loco_debug_buffer
    void* last_ws = nullptr;
    int64_t launches = 0;
    // optional per-stage CUDA-event timing (bench.py roofline): one event pair per launch
    bool prof_on = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct ProfRec { int cat; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    int prof_cat = -1;
    cudaEvent_t prof_start = nullptr;
    cudaEvent_t prof_last_end = nullptr;
};

namespace {

int fail(loco_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (cudaError_t)(expr);                                                            \
        if (e_ != cudaSuccess)                                                                           \
            return fail(h, LOCO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
    } while (0)

size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

// The stage buffers inside the workspace want 1024-byte alignment (TMA boxes, SWIZZLE_128B atoms).  The caller's pointer
// may have any alignment (cudaMalloc promises 256 B, torch's caching allocator 512 B): loco_batch_plan* report kWsSlack bytes more
// than the layout needs and the encode calls round the base up themselves.
constexpr size_t kWsSlack = 1024;
constexpr int kMaxUtts = 65535;     // utterances per call: the per-utterance kernels index them with gridDim.y / gridDim.x
bool carve_workspace(void* base, size_t bytes, size_t need, uint8_t** aligned) {
    const uintptr_t b = reinterpret_cast<uintptr_t>(base);
    const uintptr_t a = (b + 1023) & ~(uintptr_t)1023;
    const size_t lost = (size_t)(a - b);
    *aligned = reinterpret_cast<uint8_t*>(a);
    return bytes >= lost && bytes - lost >= need;
}

std::string canon_key(const char* key) {
    std::string k(key);
    const char* pres[] = {"speecht5.encoder.", "encoder."};
    for (const char* p : pres) {
        size_t n = strlen(p);
        if (k.compare(0, n, p) == 0) {
            k = k.substr(n);
            break;
        }
    }
    const std::string pc = "prenet.pos_conv_embed.conv.";
    if (k == pc + "weight_g" || k == pc + "parametrizations.weight.original0") return pc + "g";
    if (k == pc + "weight_v" || k == pc + "parametrizations.weight.original1") return pc + "v";
    return k;
}

float bf16_bits_to_float(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
float f16_bits_to_float(uint16_t hbits) {
    const uint32_t sign = (hbits >> 15) & 1, exp = (hbits >> 10) & 0x1f, man = hbits & 0x3ff;
    float v;
    if (exp == 0) v = ldexpf((float)man, -24);
    else if (exp == 31) v = man ? NAN : INFINITY;
    else v = ldexpf((float)(man | 0x400), (int)exp - 25);
    return sign ? -v : v;
}
bf16 to_bf16_host(float f) { return __float2bfloat16_rn(f); }

template <typename T>
int upload(loco_handle* h, const std::vector<T>& v, T** out) {
    void* p = nullptr;
    CK(cudaMalloc(&p, v.size() * sizeof(T)));
    h->allocs.push_back(p);
    CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<T*>(p);
    return 0;
}

int get(loco_handle* h, const std::string& key, std::initializer_list<int64_t> shape, const HostTensor** out) {
    auto it = h->host.find(key);
    if (it == h->host.end()) return fail(h, LOCO_ERR_WEIGHTS, "missing tensor: " + key);
    std::vector<int64_t> want(shape);
    if (it->second.shape != want) {
        std::string s = "bad shape for " + key + ": got [";
        for (int64_t d : it->second.shape) s += std::to_string(d) + ",";
        s += "] want [";
        for (int64_t d : want) s += std::to_string(d) + ",";
        return fail(h, LOCO_ERR_WEIGHTS, s + "]");
    }
    *out = &it->second;
    return 0;
}

int upload_f32(loco_handle* h, const std::string& key, std::initializer_list<int64_t> shape, float** out, float scale = 1.f) {
    const HostTensor* t;
    int rc = get(h, key, shape, &t);
    if (rc) return rc;
    std::vector<float> v(t->data);
    if (scale != 1.f)
        for (float& x : v) x *= scale;
    return upload(h, v, out);
}

int upload_bf16(loco_handle* h, const std::string& key, std::initializer_list<int64_t> shape, bf16** out) {
    const HostTensor* t;
    int rc = get(h, key, shape, &t);
    if (rc) return rc;
    std::vector<bf16> v(t->data.size());
    for (size_t i = 0; i < v.size(); ++i) v[i] = to_bf16_host(t->data[i]);
    return upload(h, v, out);
}

// Sinusoid table, computed the way HF builds it in fp32 (modeling_speecht5.py:305-320): f_j = exp(fl(j) * fl(-ln(1e4)/383)),
// angle = fl(p) * f_j, row = [sin | cos]; row `pad_token_id` is zero.
int build_sin_table(loco_handle* h, int rows) {
    const int H = kHidden, half = H / 2;
    std::vector<float> tab((size_t)rows * H);
    const float neg_emb = (float)(-(log(10000.0) / (double)(half - 1)));
    std::vector<float> f(half);
    for (int j = 0; j < half; ++j) f[j] = (float)exp((double)((float)j * neg_emb));
    for (int p = 0; p < rows; ++p) {
        float* r = tab.data() + (size_t)p * H;
        for (int j = 0; j < half; ++j) {
            const float ang = (float)p * f[j];
            r[j] = (float)sin((double)ang);
            r[half + j] = (float)cos((double)ang);
        }
    }
    if (h->cfg.pad_token_id >= 0 && h->cfg.pad_token_id < rows)
        memset(tab.data() + (size_t)h->cfg.pad_token_id * H, 0, H * sizeof(float));
    float* dev = nullptr;
    int rc = upload(h, tab, &dev);
    if (rc) return rc;
    h->sin_table = dev;  // an outgrown table stays in `allocs` until destroy; regrowth is rare
    h->sin_rows = rows;
    return 0;
}

void conv_frames(const loco_config& c, int n_samples, int* t) {
    int64_t cur = n_samples;
    for (int i = 0; i < c.num_conv_layers; ++i) {
        cur = cur >= c.conv_kernel[i] ? (cur - c.conv_kernel[i]) / c.conv_stride[i] + 1 : 0;
        t[i] = (int)cur;
    }
}

// SpeechT5ScaledPositionalEncoding table, computed the way HF builds it in fp32 (modeling_speecht5.py:405-411):
// div_j = exp(fl(2j) * fl(-ln(1e4)/768)), pe[p, 2j] = sin(fl(p) * div_j), pe[p, 2j+1] = cos(fl(p) * div_j).
int build_text_pe(loco_handle* h, int rows) {
    const int H = kHidden;
    std::vector<float> tab((size_t)rows * H);
    const float neg = (float)(-(log(10000.0) / (double)H));
    for (int j = 0; j < H / 2; ++j) {
        const float div = (float)exp((double)((float)(2 * j) * neg));
        for (int p = 0; p < rows; ++p) {
            const float ang = (float)p * div;
            tab[(size_t)p * H + 2 * j] = (float)sin((double)ang);
            tab[(size_t)p * H + 2 * j + 1] = (float)cos((double)ang);
        }
    }
    int rc = upload(h, tab, &h->txt_pe);
    if (rc) return rc;
    h->txt_pe_rows = rows;
    return 0;
}

// Geometry of a text batch: a row is a token, utterances back to back (no slot padding), only the transformer buffers.
int make_layout_text(loco_handle* h, const int32_t* n_tokens, int n_utts, Layout* L) {
    if (n_utts < 0 || n_utts > kMaxUtts) return fail(h, LOCO_ERR_INVALID, "n_utts must be in [0, " + std::to_string(kMaxUtts) + "]");
    L->n_utts = n_utts;
    L->meta.resize(n_utts);
    int64_t row = 0;
    for (int u = 0; u < n_utts; ++u) {
        if (n_tokens[u] <= 0) return fail(h, LOCO_ERR_INVALID, "text " + std::to_string(u) + " is empty");
        UttMeta& m = L->meta[u];
        m.sample_off = row;
        m.n_samples = n_tokens[u];
        m.t0 = n_tokens[u];
        m.t6 = n_tokens[u];
        m.row6 = (int32_t)row;
        m.slot6 = n_tokens[u];
        m.out_row = (int32_t)row;
        row += n_tokens[u];
        for (int f = 0; f < m.t6; f += 128) L->pc_tiles.push_back({m.row6 + f, f, m.t6, 0});
        if (m.t6 > L->max_t6) L->max_t6 = m.t6;
        if (m.slot6 > L->max_slot6) L->max_slot6 = m.slot6;
    }
    if (row > (int64_t)INT32_MAX / 4) return fail(h, LOCO_ERR_INVALID, "text batch too large");
    L->R6 = row;
    L->total_frames = row;
    L->total_samples = row;
    size_t p = 0;
    auto take = [&](size_t bytes) {
        size_t o = p;
        p = align_up(p + bytes);
        return o;
    };
    L->off_stats1 = take((size_t)L->R6 * 2 * kStatSlots * sizeof(float));
    L->off_stats2 = take((size_t)L->R6 * 2 * kStatSlots * sizeof(float));
    L->off_rowframe = take((size_t)L->R6 * sizeof(int32_t));
    auto add = [&](const char* name, int64_t rows, int64_t cols) {
        Buf b;
        b.rows = rows;
        b.cols = cols;
        b.off = take((size_t)rows * cols * sizeof(bf16));
        L->bufs[name] = b;
    };
    add("x", L->R6, kHidden);
    add("qkv", L->R6, 3 * kHidden);
    add("ctx", L->R6, kHidden);
    add("attn_res", L->R6, kHidden);
    add("ln1", L->R6, kHidden);
    add("mid", L->R6, kFfn);
    add("ffn_res", L->R6, kHidden);
    L->bufs["enc_in"] = L->bufs["x"];
    L->bytes = p;
    return 0;
}

int make_layout(loco_handle* h, const int32_t* n_samples, int n_utts, Layout* L) {
    if (n_utts < 0 || n_utts > kMaxUtts) return fail(h, LOCO_ERR_INVALID, "n_utts must be in [0, " + std::to_string(kMaxUtts) + "]");
    L->n_utts = n_utts;
    L->meta.resize(n_utts);
    int64_t row = 0, out_row = 0, off = 0;
    for (int u = 0; u < n_utts; ++u) {
        int t[8];
        if (n_samples[u] <= 0) return fail(h, LOCO_ERR_INVALID, "utterance " + std::to_string(u) + " is empty");
        conv_frames(h->cfg, n_samples[u], t);
        if (t[6] < 1)
            return fail(h, LOCO_ERR_INVALID,
                        "utterance " + std::to_string(u) + " is shorter than one encoder frame (" + std::to_string(n_samples[u]) +
                            " samples; minimum 400)");
        int slot = 1;
        for (int i = 0; i < 7; ++i) {
            const int sh = 6 - i;
            const int need = (t[i] + (1 << sh) - 1) >> sh;
            if (need > slot) slot = need;
        }
        UttMeta& m = L->meta[u];
        m.sample_off = off;
        m.n_samples = n_samples[u];
        m.t0 = t[0];
        m.t6 = t[6];
        m.row6 = (int32_t)row;
        m.slot6 = slot;
        m.out_row = (int32_t)out_row;
        row += slot;
        out_row += t[6];
        off += n_samples[u];
        for (int f = 0; f < t[6]; f += 128) L->pc_tiles.push_back({m.row6 + f, f, t[6], 0});
        if (t[0] > L->max_t0) L->max_t0 = t[0];
        if (t[6] > L->max_t6) L->max_t6 = t[6];
        if (slot > L->max_slot6) L->max_slot6 = slot;
    }
    if ((row << 6) > (int64_t)INT32_MAX / 2) return fail(h, LOCO_ERR_INVALID, "batch too large: more than 2^30 conv0 frames");
    L->c0_tile_start.resize((size_t)n_utts + 1);
    L->c0_tile_start[0] = 0;
    for (int u = 0; u < n_utts; ++u) L->c0_tile_start[u + 1] = L->c0_tile_start[u] + (L->meta[u].slot6 + 1) / 2;     // slot6 * 64 frames in 128-frame tiles
    {
        // positional conv timeline: utterances back to back with kPosPPHalo zero frames between neighbours
        const int64_t total = out_row + (int64_t)n_utts * kPosPPHalo;
        L->n_vtiles = (int)((total + kPosPPTile - 1) / kPosPPTile);
        L->pp_map.assign((size_t)L->n_vtiles * kPosPPTile + 2 * kPosPPHalo, -1);
        size_t pos = kPosPPHalo;
        for (int u = 0; u < n_utts; ++u) {
            const UttMeta& m = L->meta[u];
            for (int f = 0; f < m.t6; ++f) L->pp_map[pos + f] = m.row6 + f;
            pos += (size_t)m.t6 + kPosPPHalo;
        }
    }
    L->R6 = row;
    L->total_frames = out_row;
    L->total_samples = off;
    L->chunks = wave_stats_chunks(L->max_t0);

    size_t p = 0;
    auto take = [&](size_t bytes) {
        size_t o = p;
        p = align_up(p + bytes);
        return o;
    };
    L->off_partial = take((size_t)n_utts * L->chunks * 65 * sizeof(double));
    L->off_scale = take((size_t)n_utts * kConvDim * sizeof(float));
    L->off_shift = take((size_t)n_utts * kConvDim * sizeof(float));
    L->off_wfold = take((size_t)n_utts * kConv0FoldBytes);
    L->off_stats1 = take((size_t)L->R6 * 2 * kStatSlots * sizeof(float));
    L->off_stats2 = take((size_t)L->R6 * 2 * kStatSlots * sizeof(float));
    L->off_rowframe = take((size_t)L->R6 * sizeof(int32_t));
    auto add = [&](const char* name, int64_t rows, int64_t cols, int64_t pad_rows) {
        Buf b;
        b.rows = rows;
        b.cols = cols;
        b.off = take((size_t)(rows + pad_rows) * cols * sizeof(bf16));
        L->bufs[name] = b;
    };
    for (int i = 0; i < 7; ++i) {
        char nm[16];
        snprintf(nm, sizeof nm, "conv%d", i);
        add(nm, L->R6 << (6 - i), kConvDim, 8);
    }
    add("proj_ln", L->R6, kConvDim, 0);
    add("proj", L->R6, kHidden, 0);
    add("pos_conv", L->R6, kHidden, 0);
    add("x", L->R6, kHidden, 0);
    add("qkv", L->R6, 3 * kHidden, 0);
    add("ctx", L->R6, kHidden, 0);
    add("attn_res", L->R6, kHidden, 0);
    add("ln1", L->R6, kHidden, 0);
    add("mid", L->R6, kFfn, 0);
    add("ffn_res", L->R6, kHidden, 0);
    L->bufs["enc_in"] = L->bufs["x"];
    L->bytes = p;
    return 0;
}

enum ProfCat { CAT_GEMM = 0, CAT_ATTENTION = 1, CAT_POSCONV = 2, CAT_FRONTEND = 3, CAT_ROWOPS = 4, CAT_COUNT = 5 };

cudaEvent_t prof_event(loco_handle* h) {
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        h->ev_pool.push_back(e);
    }
    return h->ev_pool[h->ev_used++];
}
// One event per launch boundary: the event that ends launch i is the start of launch i+1 when nothing else was enqueued in
// between (prof_break() is called after memcpy / memset nodes), which halves the event traffic in the timed region.
void prof_begin(loco_handle* h, int cat, cudaStream_t s) {
    if (!h->prof_on) return;
    h->prof_cat = cat;
    if (h->prof_last_end) {
        h->prof_start = h->prof_last_end;
        return;
    }
    h->prof_start = prof_event(h);
    if (h->prof_start) cudaEventRecord(h->prof_start, s);
}
void prof_end(loco_handle* h, cudaStream_t s) {
    if (!h->prof_on || !h->prof_start) return;
    cudaEvent_t b = prof_event(h);
    if (b) {
        cudaEventRecord(b, s);
        h->prof.push_back({h->prof_cat, h->prof_start, b});
    }
    h->prof_last_end = b;
    h->prof_start = nullptr;
}
void prof_break(loco_handle* h) { h->prof_last_end = nullptr; }

int run_gemm(loco_handle* h, const GemmArgs& g, cudaStream_t s) {
    prof_begin(h, CAT_GEMM, s);
#ifdef LOCO_DEBUG
    int rc = h->gemm_impl == 1 ? gemm_simt_launch(g, s) : h->gemm_impl == 2 ? gemm_tc2_launch(g, h->num_sms, s) : gemm_tc_launch(g, h->num_sms, s);
#else
    int rc = gemm_tc2_launch(g, h->num_sms, s);
#endif
    prof_end(h, s);
    h->launches += 1;
    if (rc) return fail(h, LOCO_ERR_CUDA, std::string("gemm launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    return 0;
}

#define LAUNCH(cat, expr, n)                                                                                   \
    do {                                                                                                       \
        prof_begin(h, cat, s);                                                                                 \
        int rc_ = (expr);                                                                                      \
        prof_end(h, s);                                                                                        \
        h->launches += (n);                                                                                    \
        if (rc_) return fail(h, LOCO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString((cudaError_t)rc_)); \
    } while (0)

// No C++ exception may cross the C boundary (std::bad_alloc / std::length_error from the host-side vectors).
template <typename F>
int guarded(loco_handle* h, const char* what, F&& f) {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        return fail(h, LOCO_ERR_INVALID, std::string(what) + ": out of host memory");
    } catch (const std::exception& e) {
        return fail(h, LOCO_ERR_INVALID, std::string(what) + ": " + e.what());
    } catch (...) {
        return fail(h, LOCO_ERR_INVALID, std::string(what) + ": unknown C++ exception");
    }
}

void clear_plan_cache(loco_handle* h);

}  // namespace

extern "C" {

int loco_abi_version(void) { return LOCO_ABI_VERSION; }

void loco_default_config(loco_config* c) {
    memset(c, 0, sizeof(*c));
    c->hidden_size = 768;
    c->encoder_layers = 12;
    c->encoder_attention_heads = 12;
    c->encoder_ffn_dim = 3072;
    c->num_conv_layers = 7;
    const int k[7] = {10, 3, 3, 3, 3, 2, 2}, s[7] = {5, 2, 2, 2, 2, 2, 2};
    for (int i = 0; i < 7; ++i) {
        c->conv_dim[i] = 512;
        c->conv_kernel[i] = k[i];
        c->conv_stride[i] = s[i];
    }
    c->num_conv_pos_embeddings = 128;
    c->num_conv_pos_embedding_groups = 16;
    c->max_speech_positions = 4000;
    c->encoder_max_relative_position = 160;
    c->pad_token_id = 1;
    c->feat_extract_norm_is_group = 1;
    c->activation_is_gelu = 1;
    c->conv_bias = 0;
    c->layer_norm_eps = 1e-5f;
}

const char* loco_last_error(const loco_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int loco_create(const loco_config* cfg, int device, loco_handle** out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return LOCO_ERR_INVALID;
    }
    loco_config d;
    loco_default_config(&d);
    std::string bad;
    auto chk = [&](bool ok, const char* what) {
        if (!ok) bad += std::string(bad.empty() ? "" : ", ") + what;
    };
    chk(cfg->hidden_size == d.hidden_size, "hidden_size");
    chk(cfg->encoder_attention_heads == d.encoder_attention_heads, "encoder_attention_heads");
    chk(cfg->encoder_ffn_dim == d.encoder_ffn_dim, "encoder_ffn_dim");
    chk(cfg->encoder_layers >= 1 && cfg->encoder_layers <= 48, "encoder_layers");
    chk(cfg->num_conv_layers == 7, "num_conv_layers");
    for (int i = 0; i < 7; ++i)
        chk(cfg->conv_dim[i] == d.conv_dim[i] && cfg->conv_kernel[i] == d.conv_kernel[i] && cfg->conv_stride[i] == d.conv_stride[i],
            "conv_dim/kernel/stride");
    chk(cfg->num_conv_pos_embeddings == 128 && cfg->num_conv_pos_embedding_groups == 16, "num_conv_pos_embeddings/groups");
    chk(cfg->encoder_max_relative_position == 160, "encoder_max_relative_position");
    chk(cfg->feat_extract_norm_is_group == 1, "feat_extract_norm");
    chk(cfg->activation_is_gelu == 1, "activation");
    chk(cfg->conv_bias == 0, "conv_bias");
    chk(fabsf(cfg->layer_norm_eps - 1e-5f) < 1e-9f, "layer_norm_eps");
    chk(cfg->pad_token_id == 1, "pad_token_id (the sinusoid position of frame t is t + pad_token_id + 1 = t + 2, HF:349-351)");
    if (!bad.empty()) {
        g_create_error = "unsupported SpeechT5 encoder config (kernels are built for the SpeechT5-base shape family): " + bad;
        return LOCO_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = std::string("no usable CUDA device ") + std::to_string(device) + ": " +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range") +
                         " (this library has no CPU fallback)";
        return LOCO_ERR_CUDA;
    }
    cudaDeviceProp prop;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice/GetDeviceProperties: ") + cudaGetErrorString(e);
        return LOCO_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_error = "device is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; this library is sm_100a only";
        return LOCO_ERR_CUDA;
    }
    loco_handle* h = new loco_handle();
    h->cfg = *cfg;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
#ifdef LOCO_DEBUG
    if (const char* e = getenv("LOCO_ATTN_P2_MAX_FRAMES")) h->attn_p2_max_frames = atoi(e);     // A/B runs: 0 = one-item kernel only
#endif
    int rc = tensormap_init();
    if (!rc) rc = gemm_tc2_init();
    if (!rc) rc = attention_tc_init();
    if (!rc) rc = attention_p2_init();
    if (!rc) rc = posconv_pp_init();
    if (!rc) rc = conv0_tc_init();
#ifdef LOCO_DEBUG
    if (!rc) rc = conv0_mma_init();
    if (!rc) rc = posconv_tc_init();
    if (!rc) rc = gemm_tc_init();
    if (!rc) rc = attention_init();
    if (!rc) rc = posconv_init();
#endif
    if (rc) {
        g_create_error = std::string("kernel init failed: ") + cudaGetErrorString((cudaError_t)rc);
        delete h;
        return LOCO_ERR_CUDA;
    }
    *out = h;
    return LOCO_OK;
}

void loco_destroy(loco_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    clear_plan_cache(h);
    for (void* p : h->allocs) cudaFree(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    delete h;
}

static int load_tensor_impl(loco_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype);
int loco_load_tensor(loco_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype) {
    if (!h || !key || !data || ndim < 0 || ndim > 4 || (ndim > 0 && !shape)) return fail(h, LOCO_ERR_INVALID, "loco_load_tensor: bad argument");
    for (int i = 0; i < ndim; ++i)
        if (shape[i] < 0 || shape[i] > ((int64_t)1 << 31)) return fail(h, LOCO_ERR_INVALID, "loco_load_tensor: bad shape");
    return guarded(h, "loco_load_tensor", [&]() -> int { return load_tensor_impl(h, key, data, shape, ndim, dtype); });
}
static int load_tensor_impl(loco_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype) {
    if (h->finalized) return fail(h, LOCO_ERR_STATE, "weights already finalized");
    const std::string k = canon_key(key);
    if (k == "prenet.masked_spec_embed" || k.find("pos_sinusoidal_embed") != std::string::npos || k == "prenet.encode_positions.pe")
        return LOCO_OK;  // ignored: unused in eval / regenerated analytically
    const bool known = k.compare(0, 7, "prenet.") == 0 || k.compare(0, 16, "wrapped_encoder.") == 0;
    if (!known) return fail(h, LOCO_ERR_WEIGHTS, "unknown tensor key: " + std::string(key));
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    const int64_t n = t.numel();
    t.data.resize((size_t)n);
    switch (dtype) {
        case LOCO_F32: memcpy(t.data.data(), data, (size_t)n * 4); break;
        case LOCO_F64:
            for (int64_t i = 0; i < n; ++i) t.data[i] = (float)((const double*)data)[i];
            break;
        case LOCO_BF16:
            for (int64_t i = 0; i < n; ++i) t.data[i] = bf16_bits_to_float(((const uint16_t*)data)[i]);
            break;
        case LOCO_F16:
            for (int64_t i = 0; i < n; ++i) t.data[i] = f16_bits_to_float(((const uint16_t*)data)[i]);
            break;
        default: return fail(h, LOCO_ERR_INVALID, "loco_load_tensor: bad dtype");
    }
    h->host[k] = std::move(t);
    return LOCO_OK;
}

static int finalize_impl(loco_handle* h);
int loco_finalize_weights(loco_handle* h) {
    if (!h) return LOCO_ERR_INVALID;
    return guarded(h, "loco_finalize_weights", [&]() -> int { return finalize_impl(h); });
}
static int finalize_impl(loco_handle* h) {
    if (h->finalized) return fail(h, LOCO_ERR_STATE, "weights already finalized");
    CK(cudaSetDevice(h->device));
    int rc;
    const std::string fe = "prenet.feature_encoder.conv_layers.";
    h->has_speech = h->host.count(fe + "0.conv.weight") != 0;
    h->has_text = h->host.count("prenet.embed_tokens.weight") != 0;
    if (!h->has_speech && !h->has_text)
        return fail(h, LOCO_ERR_WEIGHTS, "no prenet weights: neither prenet.feature_encoder.* (speech) nor prenet.embed_tokens.weight (text)");
    if (h->has_text) {
        const HostTensor& e = h->host["prenet.embed_tokens.weight"];
        if (e.shape.size() != 2 || e.shape[1] != kHidden || e.shape[0] < 1)
            return fail(h, LOCO_ERR_WEIGHTS, "bad shape for prenet.embed_tokens.weight: want [vocab, 768]");
        h->txt_vocab = (int)e.shape[0];
        std::vector<float> v(e.data);
        if ((rc = upload(h, v, &h->txt_embed))) return rc;
        auto a = h->host.find("prenet.encode_positions.alpha");
        if (a == h->host.end() || a->second.numel() != 1) return fail(h, LOCO_ERR_WEIGHTS, "missing tensor: prenet.encode_positions.alpha");
        h->txt_alpha = a->second.data[0];
        if ((rc = build_text_pe(h, 1024))) return rc;
    }
    if (h->has_speech) {
    if ((rc = upload_f32(h, fe + "0.conv.weight", {512, 1, 10}, &h->w0))) return rc;
    if ((rc = upload_f32(h, fe + "0.layer_norm.weight", {512}, &h->gn_w))) return rc;
    if ((rc = upload_f32(h, fe + "0.layer_norm.bias", {512}, &h->gn_b))) return rc;
    for (int i = 1; i < 7; ++i) {
        const int k = h->cfg.conv_kernel[i];
        const HostTensor* t;
        if ((rc = get(h, fe + std::to_string(i) + ".conv.weight", {512, 512, k}, &t))) return rc;
        // [out][in][tap] -> [out][tap*512 + in]: a K-major GEMM weight whose K index matches the contiguous
        // strip of k input frames the implicit-GEMM A operand reads
        std::vector<bf16> w((size_t)512 * 512 * k);
        for (int o = 0; o < 512; ++o)
            for (int c = 0; c < 512; ++c)
                for (int j = 0; j < k; ++j) w[((size_t)o * k + j) * 512 + c] = to_bf16_host(t->data[((size_t)o * 512 + c) * k + j]);
        if ((rc = upload(h, w, &h->conv_w[i]))) return rc;
    }
    const std::string fp = "prenet.feature_projection.";
    if ((rc = upload_f32(h, fp + "layer_norm.weight", {512}, &h->pln_w))) return rc;
    if ((rc = upload_f32(h, fp + "layer_norm.bias", {512}, &h->pln_b))) return rc;
    if ((rc = upload_bf16(h, fp + "projection.weight", {768, 512}, &h->proj_w))) return rc;
    if ((rc = upload_f32(h, fp + "projection.bias", {768}, &h->proj_b))) return rc;
    {
        // weight-norm fold (dim = 2): W[o][i][j] = g[j] * v[o][i][j] / ||v[:, :, j]||   (HF:355-383)
        const HostTensor *g, *v;
        if ((rc = get(h, "prenet.pos_conv_embed.conv.g", {1, 1, 128}, &g))) return rc;
        if ((rc = get(h, "prenet.pos_conv_embed.conv.v", {768, 48, 128}, &v))) return rc;
        std::vector<double> norm(128, 0.0);
        for (size_t idx = 0; idx < v->data.size(); ++idx) norm[idx % 128] += (double)v->data[idx] * (double)v->data[idx];
        for (double& x : norm) x = sqrt(x);
        std::vector<bf16> w((size_t)16 * 128 * 48 * 48);     // [group][tap][out_local][in]
        std::vector<bf16> wt((size_t)16 * 128 * 48 * 48);    // [group][tap][in / 8][out_local][in % 8]
        std::vector<bf16> wp((size_t)16 * 6 * kPosPPTaps * 48 * 8, to_bf16_host(0.f));    // [group][in / 8][3 + tap][out_local][in % 8]
        for (int o = 0; o < 768; ++o)
            for (int c = 0; c < 48; ++c)
                for (int j = 0; j < 128; ++j) {
                    const float val = (float)((double)g->data[j] * (double)v->data[((size_t)o * 48 + c) * 128 + j] / norm[j]);
                    const size_t gj = (size_t)(o / 48) * 128 + j;
                    w[(gj * 48 + (o % 48)) * 48 + c] = to_bf16_host(val);
                    wt[((gj * 6 + c / 8) * 48 + (o % 48)) * 8 + (c % 8)] = to_bf16_host(val);
                    wp[((((size_t)(o / 48) * 6 + c / 8) * kPosPPTaps + 3 + j) * 48 + (o % 48)) * 8 + (c % 8)] = to_bf16_host(val);
                }
#ifdef LOCO_DEBUG
        if ((rc = upload(h, w, &h->pos_w))) return rc;
        if ((rc = upload(h, wt, &h->pos_w_tc))) return rc;
#endif
        if ((rc = upload(h, wp, &h->pos_w_pp))) return rc;
        if ((rc = upload_f32(h, "prenet.pos_conv_embed.conv.bias", {768}, &h->pos_b))) return rc;
    }
    }  // has_speech
    if ((rc = upload_f32(h, "wrapped_encoder.layer_norm.weight", {768}, &h->eln_w))) return rc;
    if ((rc = upload_f32(h, "wrapped_encoder.layer_norm.bias", {768}, &h->eln_b))) return rc;
    if ((rc = upload_bf16(h, "wrapped_encoder.embed_positions.pe_k.weight", {320, 64}, &h->pe_k))) return rc;
    if (make_tensor_map_bf16_sw128(&h->pe_map, h->pe_k, 64, 320, 64, 160))
        return fail(h, LOCO_ERR_CUDA, "cuTensorMapEncodeTiled failed for pe_k");
    h->layers.resize(h->cfg.encoder_layers);
    for (int l = 0; l < h->cfg.encoder_layers; ++l) {
        const std::string p = "wrapped_encoder.layers." + std::to_string(l) + ".";
        LayerW& w = h->layers[l];
        const HostTensor *q, *k, *v, *bq, *bk, *bv;
        if ((rc = get(h, p + "attention.q_proj.weight", {768, 768}, &q))) return rc;
        if ((rc = get(h, p + "attention.k_proj.weight", {768, 768}, &k))) return rc;
        if ((rc = get(h, p + "attention.v_proj.weight", {768, 768}, &v))) return rc;
        if ((rc = get(h, p + "attention.q_proj.bias", {768}, &bq))) return rc;
        if ((rc = get(h, p + "attention.k_proj.bias", {768}, &bk))) return rc;
        if ((rc = get(h, p + "attention.v_proj.bias", {768}, &bv))) return rc;
        // fused QKV; q (weight and bias) pre-scaled by head_dim^-0.5 (HF:891) times log2(e), so the attention
        // kernel's scores (q.k and q.pe_k are both linear in q) come out in log2 units and its softmax is a bare ex2
        const float kQScale = 0.125f * 1.4426950408889634f;
        std::vector<bf16> wqkv((size_t)2304 * 768);
        std::vector<float> bqkv(2304);
        for (size_t i = 0; i < (size_t)768 * 768; ++i) {
            wqkv[i] = to_bf16_host(q->data[i] * kQScale);
            wqkv[(size_t)768 * 768 + i] = to_bf16_host(k->data[i]);
            wqkv[(size_t)2 * 768 * 768 + i] = to_bf16_host(v->data[i]);
        }
        for (int i = 0; i < 768; ++i) {
            bqkv[i] = bq->data[i] * kQScale;
            bqkv[768 + i] = bk->data[i];
            bqkv[1536 + i] = bv->data[i];
        }
        if ((rc = upload(h, wqkv, &w.wqkv))) return rc;
        if ((rc = upload(h, bqkv, &w.bqkv))) return rc;
        if ((rc = upload_bf16(h, p + "attention.out_proj.weight", {768, 768}, &w.wo))) return rc;
        if ((rc = upload_f32(h, p + "attention.out_proj.bias", {768}, &w.bo))) return rc;
        if ((rc = upload_f32(h, p + "layer_norm.weight", {768}, &w.ln1_w))) return rc;
        if ((rc = upload_f32(h, p + "layer_norm.bias", {768}, &w.ln1_b))) return rc;
        if ((rc = upload_bf16(h, p + "feed_forward.intermediate_dense.weight", {3072, 768}, &w.w1))) return rc;
        if ((rc = upload_f32(h, p + "feed_forward.intermediate_dense.bias", {3072}, &w.b1))) return rc;
        if ((rc = upload_bf16(h, p + "feed_forward.output_dense.weight", {768, 3072}, &w.w2))) return rc;
        if ((rc = upload_f32(h, p + "feed_forward.output_dense.bias", {768}, &w.b2))) return rc;
        if ((rc = upload_f32(h, p + "final_layer_norm.weight", {768}, &w.ln2_w))) return rc;
        if ((rc = upload_f32(h, p + "final_layer_norm.bias", {768}, &w.ln2_b))) return rc;
        // ---- deferred LayerNorm operands: W' = gamma (.) W in bf16, c1 = row sums of W', c2 = W beta + b -----------------
        auto fold = [&](const float* W, const float* b, const float* gamma, const float* beta, int N, float w_scale_rows_lt, int n_scaled,
                        bf16** w_out, float** c1_out, float** c2_out) -> int {
            std::vector<bf16> wf((size_t)N * 768);
            std::vector<float> c1(N), c2(N);
            for (int n = 0; n < N; ++n) {
                const float sc = n < n_scaled ? w_scale_rows_lt : 1.0f;
                double s1 = 0.0, s2 = 0.0;
                for (int k = 0; k < 768; ++k) {
                    const float wv = W[(size_t)n * 768 + k] * sc;
                    const bf16 r = to_bf16_host(wv * gamma[k]);
                    wf[(size_t)n * 768 + k] = r;
                    s1 += (double)__bfloat162float(r);
                    s2 += (double)wv * (double)beta[k];
                }
                c1[n] = (float)s1;
                c2[n] = (float)(s2 + (double)b[n] * sc);
            }
            int rc2;
            if ((rc2 = upload(h, wf, w_out))) return rc2;
            if ((rc2 = upload(h, c1, c1_out))) return rc2;
            return upload(h, c2, c2_out);
        };
        {
            const HostTensor *w1t, *b1t, *g1, *be1;
            if ((rc = get(h, p + "feed_forward.intermediate_dense.weight", {3072, 768}, &w1t))) return rc;
            if ((rc = get(h, p + "feed_forward.intermediate_dense.bias", {3072}, &b1t))) return rc;
            if ((rc = get(h, p + "layer_norm.weight", {768}, &g1))) return rc;
            if ((rc = get(h, p + "layer_norm.bias", {768}, &be1))) return rc;
            if ((rc = fold(w1t->data.data(), b1t->data.data(), g1->data.data(), be1->data.data(), 3072, 1.0f, 0, &w.w1_ln, &w.ffn_c1, &w.ffn_c2)))
                return rc;
            const HostTensor* b2t;
            if ((rc = get(h, p + "feed_forward.output_dense.bias", {768}, &b2t))) return rc;
            std::vector<float> bb(768);
            for (int i = 0; i < 768; ++i) bb[i] = b2t->data[i] + be1->data[i];
            if ((rc = upload(h, bb, &w.b2_ln))) return rc;
        }
        if (l > 0) {
            const std::string pp = "wrapped_encoder.layers." + std::to_string(l - 1) + ".";
            const HostTensor *g2, *be2;
            if ((rc = get(h, pp + "final_layer_norm.weight", {768}, &g2))) return rc;
            if ((rc = get(h, pp + "final_layer_norm.bias", {768}, &be2))) return rc;
            std::vector<float> wcat((size_t)2304 * 768), bcat(2304);
            memcpy(wcat.data(), q->data.data(), (size_t)768 * 768 * sizeof(float));
            memcpy(wcat.data() + (size_t)768 * 768, k->data.data(), (size_t)768 * 768 * sizeof(float));
            memcpy(wcat.data() + (size_t)2 * 768 * 768, v->data.data(), (size_t)768 * 768 * sizeof(float));
            memcpy(bcat.data(), bq->data.data(), 768 * sizeof(float));
            memcpy(bcat.data() + 768, bk->data.data(), 768 * sizeof(float));
            memcpy(bcat.data() + 1536, bv->data.data(), 768 * sizeof(float));
            if ((rc = fold(wcat.data(), bcat.data(), g2->data.data(), be2->data.data(), 2304, kQScale, 768, &w.wqkv_ln, &w.qkv_c1, &w.qkv_c2)))
                return rc;
            const HostTensor* bot;
            if ((rc = get(h, p + "attention.out_proj.bias", {768}, &bot))) return rc;
            std::vector<float> bb(768);
            for (int i = 0; i < 768; ++i) bb[i] = bot->data[i] + be2->data[i];
            if ((rc = upload(h, bb, &w.bo_ln))) return rc;
        }
    }
    if (h->has_speech && (rc = build_sin_table(h, h->cfg.max_speech_positions + h->cfg.pad_token_id + 3))) return rc;
    h->host.clear();
    h->finalized = true;
    return LOCO_OK;
}

}  // extern "C"

// ---- plans -----------------------------------------------------------------------------------------------------------
namespace {

uint64_t knob_state(const loco_handle* h) {
    return ((uint64_t)(h->attn_p2 ? 1 : 0) << 60) ^ ((uint64_t)(h->attn_p2_tail ? 1 : 0) << 61) ^ ((uint64_t)(uint32_t)h->attn_p2_max_frames << 20) ^ ((uint64_t)(uint32_t)(h->attn_impl + 1) << 48) ^ ((uint64_t)(uint32_t)h->attn_tc_min_frames << 32) ^
           ((uint64_t)(uint32_t)h->attn_tc_lo << 16) ^ (uint64_t)(uint32_t)h->attn_tc_hi;
}

uint64_t hash_lengths(int kind, const int32_t* v, int n) {
    uint64_t x = 1469598103934665603ull ^ (uint64_t)kind;
    for (int i = 0; i < n; ++i) {
        x ^= (uint64_t)(uint32_t)v[i];
        x *= 1099511628211ull;
    }
    return x ^ ((uint64_t)n << 40);
}

void free_plan(loco_batch_plan* p) {
    if (!p) return;
    if (p->dev || p->uploaded) cudaSetDevice(p->device);
    if (p->chunk) --p->chunk->refs;           // the chunk is recycled (after a device synchronisation) once no plan points into it
    else if (p->dev) cudaFree(p->dev);        // waits for work that may still read the block
    if (p->uploaded) cudaEventDestroy(p->uploaded);
    delete p;
}

// A block of `bytes` in the plan arena: bump allocation in the newest chunk, else an empty older chunk (all its plans gone; the
// device is synchronised first, work of those plans may still be running), else a new chunk.
cudaError_t plan_arena_alloc(loco_handle* h, size_t bytes, PlanChunk** chunk, size_t* off) {
    bytes = (bytes + 255) / 256 * 256;
    PlanChunk* c = h->plan_chunks.empty() ? nullptr : h->plan_chunks.back();
    if (!c || c->used + bytes > c->cap) {
        c = nullptr;
        for (PlanChunk* o : h->plan_chunks)
            if (o->refs == 0 && o->cap >= bytes) {
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) return e;
                o->used = 0;
                c = o;
                break;
            }
        if (c) {        // becomes the newest chunk
            h->plan_chunks.erase(std::find(h->plan_chunks.begin(), h->plan_chunks.end(), c));
            h->plan_chunks.push_back(c);
        }
    }
    if (!c) {
        c = new PlanChunk();
        c->cap = bytes > kPlanChunkBytes ? bytes : kPlanChunkBytes;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->dev), c->cap);
        if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&c->host), c->cap);
        if (e != cudaSuccess) {
            if (c->dev) cudaFree(c->dev);
            delete c;
            return e;
        }
        h->plan_chunks.push_back(c);
    }
    *chunk = c;
    *off = c->used;
    c->used += bytes;
    ++c->refs;
    return cudaSuccess;
}

// Geometry + work lists + the device block.  `async_stream` null: synchronous upload (loco_plan_create; plan outside captures
// and hot loops).  Otherwise the block is copied from a pinned staging buffer on that stream, ordered before the encode that
// follows on it -- the device is never synchronised, so a queue of first-time encodes keeps the GPU busy.
int build_plan(loco_handle* h, int kind, const int32_t* lengths, int n_utts, loco_batch_plan** out, bool async = false,
               cudaStream_t async_stream = nullptr) {
    if (!h->finalized) return fail(h, LOCO_ERR_STATE, "plan before loco_finalize_weights");
    if (kind == 0 && !h->has_speech) return fail(h, LOCO_ERR_STATE, "this handle was loaded without the speech prenet (text-only weights)");
    if (kind == 1 && !h->has_text) return fail(h, LOCO_ERR_STATE, "this handle was loaded without the text prenet (prenet.embed_tokens.weight)");
    loco_batch_plan* p = new loco_batch_plan();
    p->kind = kind;
    p->device = h->device;
    p->lengths.assign(lengths, lengths + n_utts);
    p->knobs = knob_state(h);
    int rc = kind == 0 ? make_layout(h, lengths, n_utts, &p->L) : make_layout_text(h, lengths, n_utts, &p->L);
    if (rc) {
        delete p;
        return rc;
    }
    Layout& L = p->L;
    // ---- attention work list: every utterance's 128-query tiles go to the tcgen05 kernel.  (LOCO_DEBUG builds can route
    // utterances to the mma.sync cross-check kernel, chosen PER UTTERANCE by its own frame count, never by its batch-mates.)
    for (int u = 0; u < n_utts; ++u) {
        const int t6 = L.meta[u].t6;
        bool tc = true;
#ifdef LOCO_DEBUG
        tc = h->attn_impl == 0 || (h->attn_impl < 0 && (t6 >= h->attn_tc_min_frames || (t6 >= h->attn_tc_lo && t6 <= h->attn_tc_hi)));
#endif
        if (tc) {
            // Which tcgen05 kernel takes which query tile.
            // r03g (finer sweep, both kernels after setmaxnreg / the cubic exp2): the cost is a sawtooth in the frame count -- the
            // one-item kernel pays a whole 128-query tile for a tail of a few rows (128 -> 132 frames: 0.21 -> 0.47 ms), the
            // two-pipeline kernel a 64-query item (0.20 -> 0.33).  So: up to 192 frames everything goes to the two-pipeline kernel;
            // above, full 128-query tiles and tails of more than 64 rows go to the one-item kernel and a tail of at most 64 rows
            // becomes one two-pipeline item (264 frames: 0.44-0.49 either kernel alone).  Which kernel computes a query row depends
            // only on (frame count, row index).
            if (h->attn_p2 && t6 < h->attn_p2_max_frames) {
                for (int f = 0; f < t6; f += 64) p->at_tiles64.push_back({L.meta[u].row6 + f, f, t6, 0});
            } else {
                int f = 0;
                for (; t6 - f > (h->attn_p2 && h->attn_p2_tail ? 64 : 0); f += 128) p->at_tiles.push_back({L.meta[u].row6 + f, f, t6, 0});
                if (f < t6) p->at_tiles64.push_back({L.meta[u].row6 + f, f, t6, 0});
            }
        } else {
            p->at_utts.push_back(u);
            if (t6 > p->at_ms_max_t6) p->at_ms_max_t6 = t6;
        }
    }
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess) {
        // position tables grow on demand, as HF's do (HF:331-333); done here so that encoding never synchronises
        if (kind == 0 && L.max_t6 + h->cfg.pad_token_id + 1 >= h->sin_rows) {
            e = cudaDeviceSynchronize();
            if (e == cudaSuccess && (rc = build_sin_table(h, L.max_t6 + h->cfg.pad_token_id + 1024))) {
                delete p;
                return rc;
            }
        }
        if (kind == 1 && L.max_t6 > h->txt_pe_rows) {      // HF's table stops at max_text_positions (450); this one grows
            e = cudaDeviceSynchronize();
            if (e == cudaSuccess && (rc = build_text_pe(h, L.max_t6 + 1024))) {
                delete p;
                return rc;
            }
        }
    }
    auto place = [&](size_t bytes) {
        size_t o = p->dev_bytes;
        p->dev_bytes = (p->dev_bytes + bytes + 255) / 256 * 256;
        return o;
    };
    p->d_meta = place((size_t)n_utts * sizeof(UttMeta));
    p->d_pctiles = place(L.pc_tiles.size() * sizeof(PcTile));
    p->d_attiles = place(p->at_tiles.size() * sizeof(PcTile));
    p->d_attiles64 = place(p->at_tiles64.size() * sizeof(PcTile));
    p->d_atutts = place(p->at_utts.size() * sizeof(int32_t));
    p->d_ppmap = place(L.pp_map.size() * sizeof(int32_t));
    p->d_c0tiles = place(L.c0_tile_start.size() * sizeof(int32_t));
    if (p->dev_bytes == 0) p->dev_bytes = 256;
    if (async) {
        size_t off = 0;
        if (e == cudaSuccess) e = plan_arena_alloc(h, p->dev_bytes, &p->chunk, &off);
        if (e == cudaSuccess) {
            p->dev = p->chunk->dev + off;
            p->staging = p->chunk->host + off;
        }
    } else if (e == cudaSuccess) {
        e = cudaMalloc(reinterpret_cast<void**>(&p->dev), p->dev_bytes);
    }
    auto up = [&](size_t off, const void* src, size_t bytes) {
        if (e != cudaSuccess || !bytes) return;
        if (async) memcpy(p->staging + off, src, bytes);
        else e = cudaMemcpy(p->dev + off, src, bytes, cudaMemcpyHostToDevice);
    };
    up(p->d_meta, L.meta.data(), (size_t)n_utts * sizeof(UttMeta));
    up(p->d_pctiles, L.pc_tiles.data(), L.pc_tiles.size() * sizeof(PcTile));
    up(p->d_attiles, p->at_tiles.data(), p->at_tiles.size() * sizeof(PcTile));
    up(p->d_attiles64, p->at_tiles64.data(), p->at_tiles64.size() * sizeof(PcTile));
    up(p->d_atutts, p->at_utts.data(), p->at_utts.size() * sizeof(int32_t));
    up(p->d_ppmap, L.pp_map.data(), L.pp_map.size() * sizeof(int32_t));
    up(p->d_c0tiles, L.c0_tile_start.data(), L.c0_tile_start.size() * sizeof(int32_t));
    if (async) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->dev, p->staging, p->dev_bytes, cudaMemcpyHostToDevice, async_stream);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->uploaded, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(p->uploaded, async_stream);
        p->upload_stream = async_stream;
    }
    if (e != cudaSuccess) {
        free_plan(p);
        return fail(h, LOCO_ERR_CUDA, std::string("loco_plan_create: ") + cudaGetErrorString(e));
    }
    *out = p;
    return LOCO_OK;
}

constexpr size_t kPlanCacheSize = 256;

void clear_plan_cache(loco_handle* h) {
    for (loco_batch_plan* p : h->plan_lru) free_plan(p);
    for (PlanChunk* c : h->plan_chunks) {       // cudaFree waits for work that may still read the blocks
        cudaFree(c->dev);
        cudaFreeHost(c->host);
        delete c;
    }
    h->plan_chunks.clear();
    h->plan_lru.clear();
    h->plan_index.clear();
    h->last_plan = nullptr;
}

// The plan of this batch, built on first sight (uploaded on the encoding stream, no synchronisation) and reused afterwards.
int cached_plan(loco_handle* h, int kind, const int32_t* lengths, int n_utts, const loco_batch_plan** out, cudaStream_t s) {
    const uint64_t key = hash_lengths(kind, lengths, n_utts);
    auto it = h->plan_index.find(key);
    if (it != h->plan_index.end()) {
        loco_batch_plan* p = *it->second;
        if (p->kind == kind && (int)p->lengths.size() == n_utts && p->knobs == knob_state(h) &&
            std::equal(lengths, lengths + n_utts, p->lengths.begin())) {
            h->plan_lru.splice(h->plan_lru.begin(), h->plan_lru, it->second);
            *out = p;
            return LOCO_OK;
        }
        if (h->last_plan == p) h->last_plan = nullptr;
        free_plan(p);                       // same hash, different batch (or stale knobs): replace
        h->plan_lru.erase(it->second);
        h->plan_index.erase(it);
    }
    loco_batch_plan* p = nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);           // a capture cannot hold the allocation anyway: the synchronous form reports it
    int rc = build_plan(h, kind, lengths, n_utts, &p, cap == cudaStreamCaptureStatusNone, s);
    if (rc) return rc;
    p->cached = true;
    h->plan_lru.push_front(p);
    h->plan_index[key] = h->plan_lru.begin();
    if (h->plan_lru.size() > kPlanCacheSize) {
        loco_batch_plan* old = h->plan_lru.back();
        h->plan_index.erase(hash_lengths(old->kind, old->lengths.data(), (int)old->lengths.size()));
        h->plan_lru.pop_back();
        if (h->last_plan == old) h->last_plan = nullptr;
        free_plan(old);                     // cudaFree waits for work that may still read the block
    }
    *out = p;
    return LOCO_OK;
}

}  // namespace

// The 12 post-LN transformer layers + final LayerNorm / masked mean-pool, shared by the speech and the text paths
// (SpeechT5Encoder.forward, HF modeling_speecht5.py:1250-1338).  Expects B("x") = encoder input after its LayerNorm.
static int run_transformer(loco_handle* h, const loco_batch_plan& P, uint8_t* ws, float* pooled_dev, float* hidden_dev, cudaStream_t s) {
    int rc;
    const Layout& L = P.L;
    const int n_utts = L.n_utts;
    auto B = [&](const char* name) { return reinterpret_cast<bf16*>(ws + L.bufs.at(name).off); };
    const UttMeta* meta = reinterpret_cast<const UttMeta*>(P.dev + P.d_meta);
    const int R6 = (int)L.R6;
    const PcTile* at_tiles_dev = reinterpret_cast<const PcTile*>(P.dev + P.d_attiles);
    const PcTile* at_tiles64_dev = reinterpret_cast<const PcTile*>(P.dev + P.d_attiles64);
#ifdef LOCO_DEBUG
    const int32_t* at_utts_dev = reinterpret_cast<const int32_t*>(P.dev + P.d_atutts);
#endif
    // slot padding rows of ctx are never written by the attention kernels; keep them finite (zero) so they stay finite
    // through every later layer -- the tcgen05 attention multiplies masked (P = 0) key rows into O, and 0 * NaN = NaN
    CK(cudaMemsetAsync(ws + L.bufs.at("ctx").off, 0, (size_t)L.R6 * kHidden * sizeof(bf16), s));
    prof_break(h);
    alignas(64) CUtensorMap qkv_map;
    if (make_tensor_map_bf16_sw128(&qkv_map, B("qkv"), 3 * kHidden, (uint64_t)L.R6, 3 * kHidden, 32))
        return fail(h, LOCO_ERR_CUDA, "cuTensorMapEncodeTiled failed for qkv");

    // ---- transformer layers (post-LN) -------------------------------------------------------------------
    // ln_impl 0 (default): no LayerNorm kernel inside the stack.  out_proj / FFN2 write the un-normalised sums u1 = x + attn(x)
    // ("attn_res") and u2 = x' + ffn(x') ("ffn_res") together with per-row statistics; FFN1 and the next layer's QKV consume them
    // through gamma-folded weights (EPI_LN_*), and the residual reads normalise on the fly (EPI_BIAS_LNRESIDUAL_STATS).  The last
    // layer's u2 goes to final_ln_pool, which has always normalised it itself.  ln_impl 1 keeps the LayerNorm kernels (tests).
    const int n_layers = (int)h->layers.size();
    const bool defer = h->ln_impl == 0 && h->gemm_impl == 2;
    float* stats1 = reinterpret_cast<float*>(ws + L.off_stats1);
    float* stats2 = reinterpret_cast<float*>(ws + L.off_stats2);
    for (int l = 0; l < n_layers; ++l) {
        const LayerW& w = h->layers[l];
        const bool ln_in = defer && l > 0;         // this layer's input is the previous layer's un-normalised ffn_res
        const LayerW& wp = h->layers[l > 0 ? l - 1 : 0];
        GemmArgs g = {};
        g.A = ln_in ? B("ffn_res") : B("x"); g.lda = kHidden; g.a_rows_alloc = R6; g.C = B("qkv"); g.ldc = 3 * kHidden;
        g.M = R6; g.N = 3 * kHidden; g.K = kHidden;
        if (ln_in) {
            g.W = w.wqkv_ln; g.bias = w.qkv_c2; g.c1 = w.qkv_c1; g.stats_in = stats2; g.epilogue = EPI_LN_BIAS;
        } else {
            g.W = w.wqkv; g.bias = w.bqkv; g.epilogue = EPI_BIAS;
        }
        if ((rc = run_gemm(h, g, s))) return rc;
#ifdef LOCO_DEBUG
        if (!P.at_utts.empty())
            LAUNCH(CAT_ATTENTION, launch_attention(B("qkv"), h->pe_k, meta, at_utts_dev, (int)P.at_utts.size(), P.at_ms_max_t6, B("ctx"), s), 1);
#endif
        if (!P.at_tiles64.empty())
            LAUNCH(CAT_ATTENTION, launch_attention_p2(&qkv_map, &h->pe_map, at_tiles64_dev, (int)P.at_tiles64.size(), B("ctx"), h->num_sms, s), 1);
        if (!P.at_tiles.empty())
            LAUNCH(CAT_ATTENTION, launch_attention_tc(&qkv_map, &h->pe_map, at_tiles_dev, (int)P.at_tiles.size(), B("ctx"), h->num_sms, s), 1);
        g = GemmArgs();
        g.A = B("ctx"); g.lda = kHidden; g.a_rows_alloc = R6; g.W = w.wo; g.C = B("attn_res"); g.ldc = kHidden;
        g.bias = w.bo; g.ldr = kHidden; g.M = R6; g.N = kHidden; g.K = kHidden;
        if (ln_in) {
            g.R = B("ffn_res"); g.stats_in = stats2; g.ln_gamma = wp.ln2_w; g.bias = w.bo_ln; g.stats_out = stats1;
            g.epilogue = EPI_BIAS_LNRESIDUAL_STATS;
        } else {
            g.R = B("x"); g.stats_out = stats1; g.epilogue = defer ? EPI_BIAS_RESIDUAL_STATS : EPI_BIAS_RESIDUAL;
        }
        if ((rc = run_gemm(h, g, s))) return rc;
        if (!defer) LAUNCH(CAT_ROWOPS, launch_layernorm(B("attn_res"), B("ln1"), w.ln1_w, w.ln1_b, R6, kHidden, s), 1);
        g = GemmArgs();
        g.lda = kHidden; g.a_rows_alloc = R6; g.C = B("mid"); g.ldc = kFfn; g.M = R6; g.N = kFfn; g.K = kHidden;
        if (defer) {
            g.A = B("attn_res"); g.W = w.w1_ln; g.bias = w.ffn_c2; g.c1 = w.ffn_c1; g.stats_in = stats1; g.epilogue = EPI_LN_BIAS_GELU;
        } else {
            g.A = B("ln1"); g.W = w.w1; g.bias = w.b1; g.epilogue = EPI_BIAS_GELU;
        }
        if ((rc = run_gemm(h, g, s))) return rc;
        g = GemmArgs();
        g.A = B("mid"); g.lda = kFfn; g.a_rows_alloc = R6; g.W = w.w2; g.C = B("ffn_res"); g.ldc = kHidden;
        g.bias = w.b2; g.ldr = kHidden; g.M = R6; g.N = kHidden; g.K = kFfn;
        if (defer) {
            g.R = B("attn_res"); g.stats_in = stats1; g.ln_gamma = w.ln1_w; g.bias = w.b2_ln; g.stats_out = stats2;
            g.epilogue = EPI_BIAS_LNRESIDUAL_STATS;
        } else {
            g.R = B("ln1"); g.epilogue = EPI_BIAS_RESIDUAL;
        }
        if ((rc = run_gemm(h, g, s))) return rc;
        const bool last = (l == n_layers - 1) || (l == h->stop_after_layer);
        if (!last) {
            if (!defer) LAUNCH(CAT_ROWOPS, launch_layernorm(B("ffn_res"), B("x"), w.ln2_w, w.ln2_b, R6, kHidden, s), 1);
        } else {
            // last LayerNorm fused with the masked mean-pool (+ optional compact fp32 last_hidden_state)
            LAUNCH(CAT_ROWOPS, launch_final_ln_pool(B("ffn_res"), w.ln2_w, w.ln2_b, meta, n_utts, pooled_dev, hidden_dev, h->head, s), 1);
            break;
        }
    }
    return LOCO_OK;
}

namespace {

// Everything loco_encode / loco_encode_text / loco_encode_planned enqueue.  Kernels and memset nodes only.
int encode_with_plan(loco_handle* h, const loco_batch_plan& P, const void* input_dev, float* pooled_dev, float* hidden_dev, void* workspace_dev,
                     size_t workspace_bytes, cudaStream_t s) {
    const Layout& L = P.L;
    const int n_utts = L.n_utts;
    if (n_utts == 0) return LOCO_OK;
    if (!input_dev || !pooled_dev || !workspace_dev) return fail(h, LOCO_ERR_INVALID, "encode: null argument");
    if ((reinterpret_cast<uintptr_t>(input_dev) & 3) != 0) return fail(h, LOCO_ERR_INVALID, "the input buffer must be 4-byte aligned");
    if (P.device != h->device) return fail(h, LOCO_ERR_INVALID, "plan belongs to another device");
    uint8_t* ws = nullptr;
    if (!carve_workspace(workspace_dev, workspace_bytes, L.bytes, &ws))
        return fail(h, LOCO_ERR_WORKSPACE, "workspace too small: need " + std::to_string(L.bytes + kWsSlack) +
                                               " bytes (the plan's workspace_bytes), got " + std::to_string(workspace_bytes));
    CK(cudaSetDevice(h->device));
    if (P.uploaded && s != P.upload_stream) CK(cudaStreamWaitEvent(s, P.uploaded, 0));
    prof_break(h);
    h->last_plan = &P;
    h->last_ws = ws;
    int rc;
    auto B = [&](const char* name) { return reinterpret_cast<bf16*>(ws + L.bufs.at(name).off); };
    const UttMeta* meta = reinterpret_cast<const UttMeta*>(P.dev + P.d_meta);
    int32_t* row_frame = reinterpret_cast<int32_t*>(ws + L.off_rowframe);
    const int R6 = (int)L.R6;
    LAUNCH(CAT_ROWOPS, launch_row_frames(meta, n_utts, L.max_slot6, row_frame, s), 1);
    if (P.kind == 1) {
        LAUNCH(CAT_ROWOPS, launch_text_prenet_ln(reinterpret_cast<const int32_t*>(input_dev), h->txt_embed, h->txt_pe, h->txt_alpha, h->txt_vocab,
                                                 row_frame, B("x"), h->eln_w, h->eln_b, R6, s), 1);
        return run_transformer(h, P, ws, pooled_dev, hidden_dev, s);
    }
    const float* wave_dev = reinterpret_cast<const float*>(input_dev);
    double* partial = reinterpret_cast<double*>(ws + L.off_partial);
    float* scale = reinterpret_cast<float*>(ws + L.off_scale);
    float* shift = reinterpret_cast<float*>(ws + L.off_shift);
#ifdef LOCO_DEBUG
    const PcTile* pc_tiles = reinterpret_cast<const PcTile*>(P.dev + P.d_pctiles);
#endif
    for (int i = 0; i < 6; ++i) {  // the 8 pad frames the last implicit-GEMM rows of layer i+1 may touch
        char nm[16];
        snprintf(nm, sizeof nm, "conv%d", i);
        const Buf& b = L.bufs.at(nm);
        CK(cudaMemsetAsync(ws + b.off + (size_t)b.rows * kConvDim * 2, 0, (size_t)8 * kConvDim * 2, s));
    }
    prof_break(h);
    // ---- conv feature encoder -----------------------------------------------------------------------
    bf16* wfold = reinterpret_cast<bf16*>(ws + L.off_wfold);
    LAUNCH(CAT_FRONTEND, launch_wave_stats(wave_dev, meta, n_utts, L.chunks, h->w0, h->gn_w, h->gn_b, partial, scale, shift, wfold, s), 2);
#ifdef LOCO_DEBUG
    if (h->conv0_impl == 1)
        LAUNCH(CAT_FRONTEND, launch_conv0(wave_dev, meta, n_utts, L.max_slot6 << 6, h->w0, scale, shift, B("conv0"), s), 1);
    else
#endif
        LAUNCH(CAT_FRONTEND, launch_conv0_tc(wave_dev, meta, reinterpret_cast<const int32_t*>(P.dev + P.d_c0tiles), n_utts, L.c0_tile_start[n_utts],
                                             wfold, B("conv0"), L.R6 << 6, h->num_sms, s), 1);
    for (int i = 1; i < 7; ++i) {
        char in[16], out[16];
        snprintf(in, sizeof in, "conv%d", i - 1);
        snprintf(out, sizeof out, "conv%d", i);
        const int K = h->cfg.conv_kernel[i] * kConvDim;
        const int64_t rows_in = L.bufs.at(in).rows + 8;
        GemmArgs g = {};
        g.A = B(in);
        g.lda = 2 * kConvDim;
        g.a_rows_alloc = (rows_in * kConvDim - K) / (2 * kConvDim) + 1;
        g.W = h->conv_w[i];
        g.C = B(out);
        g.ldc = kConvDim;
        g.M = (int)L.bufs.at(out).rows;
        g.N = kConvDim;
        g.K = K;
        g.epilogue = EPI_BIAS_GELU;
        if ((rc = run_gemm(h, g, s))) return rc;
    }
    // ---- feature projection, positional conv, sinusoid, encoder input LayerNorm -------------------------
    LAUNCH(CAT_ROWOPS, launch_layernorm(B("conv6"), B("proj_ln"), h->pln_w, h->pln_b, R6, kConvDim, s), 1);
    {
        GemmArgs g = {};
        g.A = B("proj_ln"); g.lda = kConvDim; g.a_rows_alloc = R6; g.W = h->proj_w; g.C = B("proj"); g.ldc = kHidden;
        g.bias = h->proj_b; g.M = R6; g.N = kHidden; g.K = kConvDim; g.epilogue = EPI_BIAS;
        if ((rc = run_gemm(h, g, s))) return rc;
    }
#ifdef LOCO_DEBUG
    if (h->posconv_impl == 1)
        LAUNCH(CAT_POSCONV, launch_posconv(B("proj"), h->pos_w, h->pos_b, meta, n_utts, L.max_t6, B("pos_conv"), s), 1);
    else if (h->posconv_impl == 2)
        LAUNCH(CAT_POSCONV, launch_posconv_tc(B("proj"), h->pos_w_tc, h->pos_b, pc_tiles, (int)L.pc_tiles.size(), B("pos_conv"), s), 1);
    else
#endif
        LAUNCH(CAT_POSCONV, launch_posconv_pp(B("proj"), h->pos_w_pp, h->pos_b, reinterpret_cast<const int32_t*>(P.dev + P.d_ppmap), L.n_vtiles,
                                              B("pos_conv"), h->num_sms, s), 1);
    LAUNCH(CAT_ROWOPS, launch_prenet_ln(B("proj"), B("pos_conv"), h->sin_table, row_frame, B("x"), h->eln_w, h->eln_b, R6, s), 1);
    return run_transformer(h, P, ws, pooled_dev, hidden_dev, s);
}

}  // namespace

extern "C" {

int loco_plan(loco_handle* h, const int32_t* n_samples, int n_utts, int32_t* frames, int32_t* rows, int64_t* total_frames,
              size_t* workspace_bytes) {
    if (!h || (!n_samples && n_utts > 0)) return fail(h, LOCO_ERR_INVALID, "loco_plan: bad argument");
    return guarded(h, "loco_plan", [&]() -> int {
        Layout L;
        int rc = make_layout(h, n_samples, n_utts, &L);
        if (rc) return rc;
        for (int u = 0; u < n_utts; ++u) {
            if (frames) frames[u] = L.meta[u].t6;
            if (rows) rows[u] = L.meta[u].row6;
        }
        if (total_frames) *total_frames = L.total_frames;
        if (workspace_bytes) *workspace_bytes = L.bytes + kWsSlack;
        return LOCO_OK;
    });
}

int loco_plan_create(loco_handle* h, int kind, const int32_t* lengths, int n_utts, loco_batch_plan** out) {
    if (!h || !out || (!lengths && n_utts > 0) || kind < 0 || kind > 1) return fail(h, LOCO_ERR_INVALID, "loco_plan_create: bad argument");
    return guarded(h, "loco_plan_create", [&]() -> int { return build_plan(h, kind, lengths, n_utts, out); });
}

int loco_plan_info(const loco_batch_plan* p, int32_t* frames, int32_t* rows, int64_t* total_frames, size_t* workspace_bytes) {
    if (!p) return LOCO_ERR_INVALID;
    for (int u = 0; u < p->L.n_utts; ++u) {
        if (frames) frames[u] = p->L.meta[u].t6;
        if (rows) rows[u] = p->L.meta[u].row6;
    }
    if (total_frames) *total_frames = p->L.total_frames;
    if (workspace_bytes) *workspace_bytes = p->L.bytes + kWsSlack;
    return LOCO_OK;
}

void loco_plan_destroy(loco_handle* h, loco_batch_plan* p) {
    if (!p || p->cached) return;
    if (h && h->last_plan == p) h->last_plan = nullptr;
    free_plan(p);
}

int loco_encode_planned(loco_handle* h, const loco_batch_plan* plan, const void* input_dev, float* pooled_dev, float* hidden_dev,
                        void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h) return LOCO_ERR_INVALID;
    if (!plan) return fail(h, LOCO_ERR_INVALID, "loco_encode_planned: null plan");
    return guarded(h, "loco_encode_planned", [&]() -> int {
        return encode_with_plan(h, *plan, input_dev, pooled_dev, hidden_dev, workspace_dev, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
    });
}

int loco_sync_check(loco_handle* h, void* stream) {
    if (!h) return LOCO_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaGetLastError());
    return LOCO_OK;
}

int loco_encode(loco_handle* h, const float* wave_dev, const int32_t* n_samples, int n_utts, float* pooled_dev, float* hidden_dev,
                void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h) return LOCO_ERR_INVALID;
    if (!h->finalized) return fail(h, LOCO_ERR_STATE, "loco_encode before loco_finalize_weights");
    if (!h->has_speech) return fail(h, LOCO_ERR_STATE, "loco_encode: this handle was loaded without the speech prenet (text-only weights)");
    if (n_utts == 0) return LOCO_OK;
    if (!wave_dev || !n_samples || !pooled_dev || !workspace_dev) return fail(h, LOCO_ERR_INVALID, "loco_encode: null argument");
    return guarded(h, "loco_encode", [&]() -> int {
        const loco_batch_plan* plan = nullptr;
        int rc = cached_plan(h, 0, n_samples, n_utts, &plan, reinterpret_cast<cudaStream_t>(stream));
        if (rc) return rc;
        return encode_with_plan(h, *plan, wave_dev, pooled_dev, hidden_dev, workspace_dev, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
    });
}

int loco_host_workspace_bytes(loco_handle* h, const int32_t* n_samples, int n_utts, int want_hidden, size_t* bytes) {
    if (!h || !bytes || (!n_samples && n_utts > 0)) return fail(h, LOCO_ERR_INVALID, "loco_host_workspace_bytes: bad argument");
    return guarded(h, "loco_host_workspace_bytes", [&]() -> int {
        Layout L;
        int rc = make_layout(h, n_samples, n_utts, &L);
        if (rc) return rc;
        *bytes = kWsSlack + L.bytes + align_up((size_t)L.total_samples * sizeof(float)) + align_up((size_t)n_utts * kHidden * sizeof(float)) +
                 (want_hidden ? align_up((size_t)L.total_frames * kHidden * sizeof(float)) : 0);
        return LOCO_OK;
    });
}

int loco_encode_host(loco_handle* h, const float* wave_host, const int32_t* n_samples, int n_utts, float* pooled_host,
                     float* hidden_host, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h) return LOCO_ERR_INVALID;
    if (n_utts == 0) return LOCO_OK;
    if (!wave_host || !n_samples || !pooled_host || !workspace_dev) return fail(h, LOCO_ERR_INVALID, "loco_encode_host: null argument");
    if (!h->finalized) return fail(h, LOCO_ERR_STATE, "loco_encode_host before loco_finalize_weights");
    return guarded(h, "loco_encode_host", [&]() -> int {
        const loco_batch_plan* plan = nullptr;
        int rc = cached_plan(h, 0, n_samples, n_utts, &plan, reinterpret_cast<cudaStream_t>(stream));
        if (rc) return rc;
        const Layout& L = plan->L;
        const size_t wave_bytes = (size_t)L.total_samples * sizeof(float);
        const size_t pooled_bytes = (size_t)n_utts * kHidden * sizeof(float);
        const size_t hidden_bytes = hidden_host ? (size_t)L.total_frames * kHidden * sizeof(float) : 0;
        const size_t need = L.bytes + align_up(wave_bytes) + align_up(pooled_bytes) + align_up(hidden_bytes);
        uint8_t* ws = nullptr;
        if (!carve_workspace(workspace_dev, workspace_bytes, need, &ws))
            return fail(h, LOCO_ERR_WORKSPACE, "workspace too small for host encode: need " + std::to_string(need + kWsSlack) +
                                                   " bytes (loco_host_workspace_bytes), got " + std::to_string(workspace_bytes));
        CK(cudaSetDevice(h->device));
        cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
        float* wave_dev = reinterpret_cast<float*>(ws + L.bytes);
        float* pooled_dev = reinterpret_cast<float*>(ws + L.bytes + align_up(wave_bytes));
        float* hidden_dev = hidden_host ? reinterpret_cast<float*>(ws + L.bytes + align_up(wave_bytes) + align_up(pooled_bytes)) : nullptr;
        CK(cudaMemcpyAsync(wave_dev, wave_host, wave_bytes, cudaMemcpyHostToDevice, s));
        rc = encode_with_plan(h, *plan, wave_dev, pooled_dev, hidden_dev, ws, L.bytes, s);
        if (rc) return rc;
        CK(cudaMemcpyAsync(pooled_host, pooled_dev, pooled_bytes, cudaMemcpyDeviceToHost, s));
        if (hidden_host) CK(cudaMemcpyAsync(hidden_host, hidden_dev, hidden_bytes, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return LOCO_OK;
    });
}

int loco_plan_text(loco_handle* h, const int32_t* n_tokens, int n_utts, int32_t* rows, int64_t* total_tokens, size_t* workspace_bytes) {
    if (!h || (!n_tokens && n_utts > 0)) return fail(h, LOCO_ERR_INVALID, "loco_plan_text: bad argument");
    return guarded(h, "loco_plan_text", [&]() -> int {
        Layout L;
        int rc = make_layout_text(h, n_tokens, n_utts, &L);
        if (rc) return rc;
        for (int u = 0; u < n_utts; ++u)
            if (rows) rows[u] = L.meta[u].row6;
        if (total_tokens) *total_tokens = L.total_frames;
        if (workspace_bytes) *workspace_bytes = L.bytes + kWsSlack;
        return LOCO_OK;
    });
}

int loco_encode_text(loco_handle* h, const int32_t* tokens_dev, const int32_t* n_tokens, int n_utts, float* pooled_dev, float* hidden_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h) return LOCO_ERR_INVALID;
    if (!h->finalized) return fail(h, LOCO_ERR_STATE, "loco_encode_text before loco_finalize_weights");
    if (!h->has_text) return fail(h, LOCO_ERR_STATE, "loco_encode_text: this handle was loaded without the text prenet (prenet.embed_tokens.weight)");
    if (n_utts == 0) return LOCO_OK;
    if (!tokens_dev || !n_tokens || !pooled_dev || !workspace_dev) return fail(h, LOCO_ERR_INVALID, "loco_encode_text: null argument");
    return guarded(h, "loco_encode_text", [&]() -> int {
        const loco_batch_plan* plan = nullptr;
        int rc = cached_plan(h, 1, n_tokens, n_utts, &plan, reinterpret_cast<cudaStream_t>(stream));
        if (rc) return rc;
        return encode_with_plan(h, *plan, tokens_dev, pooled_dev, hidden_dev, workspace_dev, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
    });
}

int64_t loco_launch_count(const loco_handle* h) { return h ? h->launches : 0; }

int loco_profile_enable(loco_handle* h, int on) {
    if (!h) return LOCO_ERR_INVALID;
    h->prof_on = on != 0;
    h->prof.clear();
    h->ev_used = 0;
    h->prof_last_end = nullptr;
    return LOCO_OK;
}

int loco_profile_collect(loco_handle* h, int n_cats, double* ms, int64_t* launches) {
    if (!h || !ms || !launches || n_cats < CAT_COUNT) return fail(h, LOCO_ERR_INVALID, "loco_profile_collect: need room for 5 categories");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < n_cats; ++i) {
        ms[i] = 0.0;
        launches[i] = 0;
    }
    for (const auto& r : h->prof) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cat] += (double)t;
        launches[r.cat] += 1;
    }
    h->prof.clear();
    h->ev_used = 0;
    h->prof_last_end = nullptr;
    return LOCO_OK;
}

static int set_head_impl(loco_handle* h, int method, const float* q_host, const float* w_host, const float* b_host, int n_classes);
int loco_set_head(loco_handle* h, int method, const float* q_host, const float* w_host, const float* b_host, int n_classes) {
    if (!h) return LOCO_ERR_INVALID;
    return guarded(h, "loco_set_head", [&]() -> int { return set_head_impl(h, method, q_host, w_host, b_host, n_classes); });
}
static int set_head_impl(loco_handle* h, int method, const float* q_host, const float* w_host, const float* b_host, int n_classes) {
    if (method < kPoolAverage || method > kPoolAttention) return fail(h, LOCO_ERR_INVALID, "loco_set_head: method must be 0 (average), 1 (max) or 2 (self_attention)");
    if (method == kPoolAttention && !q_host) return fail(h, LOCO_ERR_INVALID, "loco_set_head: self_attention pooling needs q");
    if ((w_host == nullptr) != (b_host == nullptr) || (w_host && n_classes <= 0) || n_classes < 0)
        return fail(h, LOCO_ERR_INVALID, "loco_set_head: classifier weight, bias and n_classes go together");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());     // an encode in flight may still read the previous head
    HeadArgs a;
    a.method = method;
    a.n_classes = w_host ? n_classes : 0;
    int rc;
    if (q_host) {
        std::vector<float> q(q_host, q_host + kHidden);
        if ((rc = upload(h, q, const_cast<float**>(&a.q)))) return rc;
    }
    if (w_host) {
        std::vector<float> w(w_host, w_host + (size_t)n_classes * kHidden), b(b_host, b_host + n_classes);
        if ((rc = upload(h, w, const_cast<float**>(&a.w)))) return rc;
        if ((rc = upload(h, b, const_cast<float**>(&a.b)))) return rc;
    }
    // release the previous head's arrays (the device is idle: synchronised above)
    for (const float* old : {h->head.q, h->head.w, h->head.b}) {
        if (!old) continue;
        auto it = std::find(h->allocs.begin(), h->allocs.end(), (void*)old);
        if (it != h->allocs.end()) h->allocs.erase(it);
        cudaFree((void*)old);
    }
    h->head = a;
    h->head_set = true;
    return LOCO_OK;
}

int loco_set_head_outputs(loco_handle* h, float* head_pooled_dev, float* logits_dev) {
    if (!h) return LOCO_ERR_INVALID;
    if ((head_pooled_dev || logits_dev) && !h->head_set) return fail(h, LOCO_ERR_STATE, "loco_set_head_outputs before loco_set_head");
    if (logits_dev && !h->head.w) return fail(h, LOCO_ERR_STATE, "loco_set_head_outputs: logits requested but the head has no classifier weights");
    h->head.pooled_out = head_pooled_dev;
    h->head.logits_out = logits_dev;
    return LOCO_OK;
}

int loco_debug_set(loco_handle* h, const char* name, int64_t value) {
    if (!h || !name) return LOCO_ERR_INVALID;
#ifdef LOCO_DEBUG
    if (!strcmp(name, "gemm_impl")) h->gemm_impl = (int)value;
    else if (!strcmp(name, "posconv_impl")) h->posconv_impl = (int)value;
    else if (!strcmp(name, "conv0_impl")) h->conv0_impl = (int)value;
    else if (!strcmp(name, "ln_impl")) h->ln_impl = (int)value;
    else if (!strcmp(name, "attn_impl")) h->attn_impl = (int)value;
    else if (!strcmp(name, "attn_p2")) h->attn_p2 = value != 0;
    else if (!strcmp(name, "attn_p2_max_frames")) h->attn_p2_max_frames = (int)value;
    else if (!strcmp(name, "attn_p2_tail")) h->attn_p2_tail = value != 0;
    else if (!strcmp(name, "attn_tc_min_frames")) h->attn_tc_min_frames = (int)value;
    else if (!strcmp(name, "attn_tc_lo")) h->attn_tc_lo = (int)value;
    else if (!strcmp(name, "attn_tc_hi")) h->attn_tc_hi = (int)value;
    else if (!strcmp(name, "stop_after_layer")) h->stop_after_layer = (int)value;
    else return fail(h, LOCO_ERR_INVALID, std::string("unknown debug knob: ") + name);
    return LOCO_OK;
#else
    (void)value;
    return fail(h, LOCO_ERR_INVALID, std::string("debug knob '") + name + "': this is the product build; the cross-check kernels and "
                                         "their switches exist only in libloco_asr_debug.so (-DLOCO_DEBUG)");
#endif
}

int loco_is_debug_build(void) {
#ifdef LOCO_DEBUG
    return 1;
#else
    return 0;
#endif
}

int loco_debug_buffer(loco_handle* h, const char* name, void** dev_ptr, int64_t* n_rows, int64_t* n_cols, int* dtype) {
    if (!h || !name) return LOCO_ERR_INVALID;
    if (!h->last_ws || !h->last_plan) return fail(h, LOCO_ERR_STATE, "no encode has run yet (or its plan was destroyed)");
    auto it = h->last_plan->L.bufs.find(name);
    if (it == h->last_plan->L.bufs.end()) return fail(h, LOCO_ERR_INVALID, std::string("unknown stage buffer: ") + name);
    if (dev_ptr) *dev_ptr = reinterpret_cast<uint8_t*>(h->last_ws) + it->second.off;
    if (n_rows) *n_rows = it->second.rows;
    if (n_cols) *n_cols = it->second.cols;
    if (dtype) *dtype = it->second.dtype;
    return LOCO_OK;
}

int loco_debug_gemm(loco_handle* h, int impl, const void* a, int64_t lda, int64_t a_rows_alloc, const void* w, void* c,
                    const float* bias, const void* r, int m, int n, int k, int epilogue, void* stream) {
    if (!h) return LOCO_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    GemmArgs g = {};
    g.A = reinterpret_cast<const bf16*>(a); g.lda = lda; g.a_rows_alloc = a_rows_alloc;
    g.W = reinterpret_cast<const bf16*>(w); g.C = reinterpret_cast<bf16*>(c); g.ldc = n; g.bias = bias;
    g.R = reinterpret_cast<const bf16*>(r); g.ldr = n; g.M = m; g.N = n; g.K = k; g.epilogue = epilogue;
#ifndef LOCO_DEBUG
    if (impl != 2) return fail(h, LOCO_ERR_INVALID, "loco_debug_gemm: only the CTA-pair kernel (impl 2) exists in the product build");
#endif
    const int saved = h->gemm_impl;
    h->gemm_impl = impl;
    int rc = run_gemm(h, g, reinterpret_cast<cudaStream_t>(stream));
    h->gemm_impl = saved;
    return rc;
}

int loco_debug_gemm_ln(loco_handle* h, const void* a, const void* w, void* c, const float* bias, const void* r, int m, int n, int k,
                       int epilogue, const float* stats_in, const float* c1, const float* gamma, float* stats_out, void* stream) {
    if (!h) return LOCO_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    GemmArgs g = {};
    g.A = reinterpret_cast<const bf16*>(a); g.lda = k; g.a_rows_alloc = m;
    g.W = reinterpret_cast<const bf16*>(w); g.C = reinterpret_cast<bf16*>(c); g.ldc = n; g.bias = bias;
    g.R = reinterpret_cast<const bf16*>(r); g.ldr = n; g.M = m; g.N = n; g.K = k; g.epilogue = epilogue;
    g.stats_in = stats_in; g.c1 = c1; g.ln_gamma = gamma; g.stats_out = stats_out;
    const int saved = h->gemm_impl;
    h->gemm_impl = 2;
    int rc = run_gemm(h, g, reinterpret_cast<cudaStream_t>(stream));
    h->gemm_impl = saved;
    return rc;
}

}  // extern "C"
