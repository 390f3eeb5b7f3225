// Conv layer 0 + GroupNorm + GELU on tcgen05 (product path; frontend.cu keeps the mma.sync kernel as a debug cross-check and
// the waveform-moment kernels that produce the GroupNorm statistics).  SpeechT5GroupNormConvLayer, HF modeling_speecht5.py:260-281.
//
// The conv is a [frames, 10] x [10, 512] product.  As in the mma.sync kernel the GroupNorm scale is folded into the weights and
// fp32 accuracy comes from a 3-term bf16 split; here the split is laid along K so ONE K = 48 GEMM does it:
//     A row (frame)   = [ x_hi(10 taps) 1 0.. | x_lo(10 taps) 0.. | x_hi(10 taps) 1 0.. ]          (3 x 16)
//     B row (channel) = [ w'_hi(10)  shift_hi | w'_hi(10)      0  | w'_lo(10)  shift_lo ]
// so the GroupNorm shift rides in the padding taps (as a hi + lo pair) and the accumulator IS the normalised activation: no
// accumulator initialisation, no ldmatrix / mma.sync instruction stream.  gn_finalize_kernel writes B once per utterance
// (48 KB, already in the UMMA no-swizzle K-major layout); this kernel is then only: build A (128 threads, one frame each),
// six 128 x 256 x 16 MMAs per 128-frame tile into two 256-column accumulators, and sixteen epilogue warps
// (TMEM -> GELU -> bf16 -> swizzled smem panel -> TMA store).  The mma.sync kernel issued 57 % of its slots and kept the
// shared-memory pipe 60 % busy with B fragments and staging for a 43 % HBM write rate (ncu r03j); what remains per output here
// is the GELU itself.
//
// Persistent, one CTA per SM, each CTA a contiguous range of the batch's 128-frame tiles (an utterance's tiles are consecutive,
// so B is reloaded only when the range crosses into the next utterance).  Accumulator h (channels 256 h ..) belongs to epilogue
// group h (8 warps); a group releases its accumulator after its last tcgen05.ld, so the next tile's three MMAs run under the
// math of the last column chunk.
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int CT_ROWS = 128;                       // frames per tile (UMMA M)
constexpr int CT_KC = 6;                           // K = 48 as six 16-byte chunks
constexpr int CT_A_BYTES = CT_KC * CT_ROWS * 16;   // 12288
constexpr int CT_B_BYTES = CT_KC * kConvDim * 16;  // 49152
static_assert(CT_B_BYTES == kConv0FoldBytes, "gn_finalize_kernel writes this layout");
constexpr int CT_EPI_WARPS = 16;
constexpr int CT_PANEL_BYTES = 32 * 128;           // 32 frames x 64 channels, 128-byte swizzle
constexpr int CT_THREADS = 768;                    // warp 0 B loader, warp 1 MMA, warps 2-3 idle, 4-19 epilogue, 20-23 A builders
constexpr int CT_OFF_B = CT_EPI_WARPS * 2 * CT_PANEL_BYTES;      // 131072
constexpr int CT_OFF_A = CT_OFF_B + CT_B_BYTES;                  // 180224
constexpr int CT_OFF_BARS = CT_OFF_A + 2 * CT_A_BYTES;           // 204800
constexpr int CT_SMEM = CT_OFF_BARS + 256 + 1024;

struct __align__(8) CtBars {
    uint64_t b_full, b_empty;
    uint64_t a_full[2], a_empty[2];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// Walks this CTA's tile range as (utterance, tile inside the utterance) pairs.
struct TileWalk {
    const int32_t* tile_start;
    int g, g_end, u, u_end;      // u_end: first global tile of the next utterance
    __device__ TileWalk(const int32_t* ts, int n_utts, int g0, int g1) : tile_start(ts), g(g0), g_end(g1) {
        int lo = 0, hi = n_utts - 1;             // last u with tile_start[u] <= g0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(ts + mid) <= g0) lo = mid; else hi = mid - 1;
        }
        u = lo;
        u_end = __ldg(ts + u + 1);
        skip_empty();
    }
    __device__ void skip_empty() {
        while (g < g_end && g >= u_end) {
            ++u;
            u_end = __ldg(tile_start + u + 1);
        }
    }
    __device__ bool done() const { return g >= g_end; }
    __device__ int local_tile() const { return g - __ldg(tile_start + u); }
    __device__ bool last_of_utt() const { return g + 1 >= u_end || g + 1 >= g_end; }
    __device__ void next() {
        ++g;
        skip_empty();
    }
};


__global__ void __launch_bounds__(CT_THREADS, 1)
conv0_tc_kernel(const float* __restrict__ wave, const UttMeta* __restrict__ meta, const int32_t* __restrict__ tile_start, int n_utts,
                int n_tiles, const bf16* __restrict__ wfold, const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    CtBars* bars = reinterpret_cast<CtBars*>(smem_al + CT_OFF_BARS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        tma_prefetch_desc(&out_map);
        mbar_init(smem_u32(&bars->b_full), 1);
        mbar_init(smem_u32(&bars->b_empty), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->a_full[b]), 128);
            mbar_init(smem_u32(&bars->a_empty[b]), 1);
            mbar_init(smem_u32(&bars->acc_full[b]), 1);
            mbar_init(smem_u32(&bars->acc_empty[b]), 8 * 32);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), 512);
    pdl_launch_dependents();
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const int g0 = (int)(((int64_t)blockIdx.x * n_tiles) / gridDim.x), g1 = (int)(((int64_t)(blockIdx.x + 1) * n_tiles) / gridDim.x);

    // setmaxnreg moves registers inside the CTA's own allocation (768 x 80 at launch): the two auxiliary warpgroups drop to 48 and
    // free 2 x 128 x 32 = 8192, exactly what the four epilogue warpgroups need to rise to 96 (asking for more blocks forever)
    if (warp == 0) {
        setmaxnreg_dec<48>();
        // ===================== B loader: the utterance's folded weights, 48 KB, on every utterance change =====================
        if (lane == 0) {
            uint32_t n = 0;
            int last_u = -1;
            for (TileWalk w(tile_start, n_utts, g0, g1); !w.done(); w.next()) {
                if (w.u == last_u) continue;
                last_u = w.u;
                mbar_wait(smem_u32(&bars->b_empty), (n & 1u) ^ 1u);
                const uint32_t full = smem_u32(&bars->b_full);
                mbar_arrive_expect_tx(full, CT_B_BYTES);
                bulk_load_1d(smem_base + CT_OFF_B, reinterpret_cast<const uint8_t*>(wfold) + (size_t)w.u * CT_B_BYTES, CT_B_BYTES, full);
                ++n;
            }
        }
    } else if (warp == 1) {
        setmaxnreg_dec<48>();
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16(CT_ROWS, 256);
        const uint64_t da0 = umma_desc_noswizzle_kmajor(smem_base + CT_OFF_A, CT_ROWS * 16, 128);
        const uint64_t db0 = umma_desc_noswizzle_kmajor(smem_base + CT_OFF_B, kConvDim * 16, 128);
        constexpr uint64_t kAStepK = 2 * CT_ROWS;            // two 16-byte chunk columns, in 16 B units
        constexpr uint64_t kBStepK = 2 * kConvDim;
        uint32_t n = 0, nb = 0;
        int last_u = -1;
        for (TileWalk w(tile_start, n_utts, g0, g1); !w.done(); w.next(), ++n) {
            const uint32_t ab = n & 1u, use = (n >> 1) & 1u;
            if (w.u != last_u) {
                last_u = w.u;
                mbar_wait(smem_u32(&bars->b_full), nb & 1u);
                ++nb;
            }
            mbar_wait(smem_u32(&bars->a_full[ab]), use);
            const bool release_b = w.last_of_utt();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                mbar_wait(smem_u32(&bars->acc_empty[hh]), (n & 1u) ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = da0 + (uint64_t)(ab * (CT_A_BYTES / 16));
                    const uint64_t db = db0 + (uint64_t)(hh * 256);
                    umma_bf16(tmem_base + hh * 256, da, db, idesc, 0u);
                    umma_bf16(tmem_base + hh * 256, da + kAStepK, db + kBStepK, idesc, 1u);
                    umma_bf16(tmem_base + hh * 256, da + 2 * kAStepK, db + 2 * kBStepK, idesc, 1u);
                    umma_commit(smem_u32(&bars->acc_full[hh]));
                    if (hh == 1) {
                        umma_commit(smem_u32(&bars->a_empty[ab]));
                        if (release_b) umma_commit(smem_u32(&bars->b_empty));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= 20) {
        setmaxnreg_dec<48>();
        // ===================== A builders: one frame per thread, hi / lo split of its ten taps =====================
        const int bt = tid - 20 * 32;
        uint32_t n = 0;
        int last_u = -1;
        const float* x = nullptr;
        int n_samples = 0;
        for (TileWalk w(tile_start, n_utts, g0, g1); !w.done(); w.next(), ++n) {
            const uint32_t ab = n & 1u, use = (n >> 1) & 1u;
            if (w.u != last_u) {
                last_u = w.u;
                const UttMeta m = meta[w.u];
                x = wave + m.sample_off;
                n_samples = m.n_samples;
            }
            const int s0 = (w.local_tile() * CT_ROWS + bt) * 5;
            float v[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) v[k] = s0 + k < n_samples ? __ldg(x + s0 + k) : 0.f;
            uint32_t hi[5], lo[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const bf16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
                const bf16 l0 = __float2bfloat16_rn(v[2 * k] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v[2 * k + 1] - __bfloat162float(h1));
                hi[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lo[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            mbar_wait(smem_u32(&bars->a_empty[ab]), use ^ 1u);
            const uint32_t dst = smem_base + CT_OFF_A + ab * CT_A_BYTES + bt * 16;
            const uint4 c0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), c1 = make_uint4(hi[4], 0x3F80u, 0u, 0u);     // 0x3F80 = bf16 1.0 in tap slot 10
            sts128(dst, c0);
            sts128(dst + 1 * CT_ROWS * 16, c1);
            sts128(dst + 2 * CT_ROWS * 16, make_uint4(lo[0], lo[1], lo[2], lo[3]));
            sts128(dst + 3 * CT_ROWS * 16, make_uint4(lo[4], 0u, 0u, 0u));
            sts128(dst + 4 * CT_ROWS * 16, c0);
            sts128(dst + 5 * CT_ROWS * 16, c1);
            fence_proxy_async_smem();
            mbar_arrive(smem_u32(&bars->a_full[ab]));
        }
    } else if (warp >= 4) {
        setmaxnreg_inc<96>();
        // ===================== epilogue: group e = accumulator e = channels 256 e .. 256 e + 255 =====================
        const int ew = warp - 4, e = ew >> 3, q = warp & 3, c2 = (ew >> 2) & 1;
        const uint32_t panel0 = smem_base + ew * 2 * CT_PANEL_BYTES;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(e * 256 + c2 * 128);
        const int col0 = e * 256 + c2 * 128;
        uint32_t n = 0;
        int last_u = -1, t0 = 0, slot0 = 0, row0 = 0;
        for (TileWalk w(tile_start, n_utts, g0, g1); !w.done(); w.next(), ++n) {
            if (w.u != last_u) {
                last_u = w.u;
                const UttMeta m = meta[w.u];
                t0 = m.t0;
                slot0 = m.slot6 << 6;
                row0 = m.row6 << 6;
            }
            const int f_box = w.local_tile() * CT_ROWS + q * 32;
            const bool box_ok = f_box < slot0;          // slots are multiples of 64 frames: a 32-frame box is inside or outside
            const bool row_ok = f_box + lane < t0;      // slot padding frames are written as zeros
            mbar_wait(smem_u32(&bars->acc_full[e]), n & 1u);
            tc_fence_after();
            uint32_t v[2][32];
            tmem_ld_32x32(t_row, v[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t panel = panel0 + (c >> 1) * CT_PANEL_BYTES;
                const uint32_t my_row = panel + lane * 128;
                if ((c & 1) == 0) {
                    if (lane == 0) bulk_wait_read<1>();      // the store that last read this panel is done (the other panel's may be in flight)
                    __syncwarp();
                }
                tmem_ld_wait(v[c & 1]);
                if (c + 1 < 4) {
                    tmem_ld_32x32(t_row + (c + 1) * 32, v[(c + 1) & 1]);
                } else {
                    tc_fence_before();
                    mbar_arrive(smem_u32(&bars->acc_empty[e]));
                }
                if (box_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 o = make_uint4(0u, 0u, 0u, 0u);
                        if (row_ok) {
                            float2 f[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                f[i] = gelu_erf2(make_float2(__uint_as_float(v[c & 1][j + 2 * i]), __uint_as_float(v[c & 1][j + 2 * i + 1])));
                            o.x = pack_bf16(f[0].x, f[0].y);
                            o.y = pack_bf16(f[1].x, f[1].y);
                            o.z = pack_bf16(f[2].x, f[2].y);
                            o.w = pack_bf16(f[3].x, f[3].y);
                        }
                        sts128(my_row + ((((c & 1) * 4 + (j >> 3)) ^ (lane & 7)) << 4), o);
                    }
                    if (c & 1) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&out_map, panel, col0 + (c >> 1) * 64, row0 + f_box);
                            bulk_commit();
                        }
                    }
                }
            }
        }
        if (lane == 0) bulk_wait<0>();
    } else {
        setmaxnreg_dec<48>();       // warps 2-3: the rest of warpgroup 0
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int conv0_tc_init() {
    return (int)cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM);
}

int launch_conv0_tc(const float* wave, const UttMeta* meta, const int32_t* tile_start, int n_utts, int n_tiles, const bf16* wfold,
                    bf16* out, int64_t out_rows, int num_sms, cudaStream_t s) {
    if (n_utts <= 0 || n_tiles <= 0) return 0;
    alignas(64) CUtensorMap map;
    int rc = make_tensor_map_bf16_sw128(&map, out, (uint64_t)kConvDim, (uint64_t)out_rows, (uint64_t)kConvDim, 32);
    if (rc) return rc;
    const int grid = n_tiles < num_sms ? n_tiles : num_sms;
    return launch_pdl(conv0_tc_kernel, dim3(grid), dim3(CT_THREADS), (size_t)CT_SMEM, s, wave, meta, tile_start, n_utts, n_tiles, wfold, map);
}

}  // namespace loco
