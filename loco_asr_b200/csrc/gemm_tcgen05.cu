// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma with
// the fp32 accumulator in TMEM -> tcgen05.ld epilogue (bias / exact GELU / residual) -> swizzled smem panel ->
// TMA store (the residual operand arrives through the same panel by TMA load).
//
// One kernel covers every GEMM-shaped stage of the encoder:
//   a3  conv layers 1-6 as implicit GEMM (A rows overlap: lda = 2*512, K = k*512)   HF modeling_speecht5.py:210-228
//   a4  feature projection 512 -> 768                                               HF:498-510
//   a12 fused QKV projection 768 -> 2304 and out_proj (+ residual)                  HF:872-986
//   a13 FFN 768 -> 3072 (+ GELU) and 3072 -> 768 (+ residual)                       HF:989-1010
//
// Tile: 128 (M) x 256 (N) x 64 (K) per stage, 4 smem stages (4 x 48 KB), two 256-column TMEM accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1.  Roles: warp 0 = TMA producer (one lane),
// warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator, warps 4-11 = epilogue: warp w owns TMEM lanes
// 32*(w%4)..+31 (one accumulator row per thread) and column half (w-4)/4, i.e. two warps per SM sub-partition so
// the epilogue's dependent FP32/MUFU chains overlap (with one warp per sub-partition the GELU epilogue, not
// the MMA, set the pace: ncu r1a, tensor pipe 14 % active on FFN1).  Each epilogue warp owns a 4 KB SWIZZLE_128B
// staging panel (32 rows x 64 columns): with one accumulator row per thread, direct global stores / residual loads
// touch 32 different 128-byte lines per instruction and made the K = 768 GEMMs LSU-bound (out_proj 48 % of peak);
// through the panel every global access is a full-line TMA transfer and the M tail is clipped by the tensor map.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
// 4 smem stages and one staging panel per epilogue warp.  (A 3-stage / two-panel variant that prefetches both residual
// panels of a tile while its MMAs run was measured and lost: out_proj 977 -> 896, FFN2 1278 -> 1185 TFLOP/s in
// tools/gemm_sweep.py -- the fourth operand stage is worth more than hiding the residual round trip.)
template <int EPI> struct GemmCfg {
    static constexpr int STAGES = 4;
    static constexpr int PANELS = 1;
};
constexpr int MAX_STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int PANEL_BYTES = 32 * 128;       // one epilogue warp's staging panel: 32 rows x 64 bf16
template <int EPI> constexpr int smem_bytes() {
    return GemmCfg<EPI>::STAGES * STAGE_BYTES + GemmCfg<EPI>::PANELS * NUM_EPI_WARPS * PANEL_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
}

struct __align__(8) Barriers {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tmem_full[ACC_STAGES];
    uint64_t tmem_empty[ACC_STAGES];
    uint64_t res_full[NUM_EPI_WARPS];   // residual panel landed (one per epilogue warp)
    uint32_t tmem_base;
};

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r,
               const float* __restrict__ bias, int M, int N, int K) {
    constexpr int STAGES = GemmCfg<EPI>::STAGES, PANELS = GemmCfg<EPI>::PANELS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
    Barriers* bars = reinterpret_cast<Barriers*>(smem_aligned + STAGES * STAGE_BYTES + PANELS * NUM_EPI_WARPS * PANEL_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles_n = N / BN;
    const int n_tiles_m = (M + BM - 1) / BM;
    const int n_tiles = n_tiles_m * n_tiles_n;
    const int n_kb = K / BK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_c);
        if (EPI == EPI_BIAS_RESIDUAL) tma_prefetch_desc(&tma_r);
        for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(smem_u32(&bars->res_full[w]), 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(smem_u32(&bars->tmem_full[a]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[a]), NUM_EPI_WARPS * 32);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 2) tmem_alloc(smem_u32(&bars->tmem_base), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();
    pdl_wait();          // barrier init / TMEM allocation / tensor-map prefetch above overlap the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * BM;
                const int n0 = (tile % n_tiles_n) * BN;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1u);
                    const uint32_t full = smem_u32(&bars->full[stage]);
                    mbar_arrive_expect_tx(full, STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    tma_load_2d(sa, &tma_a, full, kb * BK, m0);
                    tma_load_2d(sa + A_STAGE_BYTES, &tma_b, full, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(smem_u32(&bars->full[stage]), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint64_t da = umma_desc_sw128_kmajor(sa);
                    const uint64_t db = umma_desc_sw128_kmajor(sa + A_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 B per K=16 step inside the 128 B swizzle atom (descriptor start is in 16 B units)
                        umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&bars->empty[stage]));  // frees the smem stage when these MMAs retire
                    if (kb == n_kb - 1) umma_commit(smem_u32(&bars->tmem_full[acc]));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> swizzled smem panel -> TMA store ==============
        const int q = warp & 3;             // TMEM lane quadrant this warp may access
        const int half = (warp - 4) >> 2;   // which 128 of the tile's 256 columns
        const uint32_t panel0 = smem_base + STAGES * STAGE_BYTES + (warp - 4) * PANELS * PANEL_BYTES;
        const uint32_t res_bar = smem_u32(&bars->res_full[warp - 4]);
        uint32_t res_phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles_n) * BM + q * 32;
            const int n0 = (tile % n_tiles_n) * BN + half * (BN / 2);
            if (EPI == EPI_BIAS_RESIDUAL && PANELS == 2) {
                // both residual panels of this warp's 32 x 128 slice, fetched while the tile's MMAs still run
                if (lane == 0) {
                    bulk_wait_read<0>();            // the previous tile's stores have finished reading the panels
                    mbar_arrive_expect_tx(res_bar, 2 * PANEL_BYTES);
                    tma_load_2d(panel0, &tma_r, res_bar, n0, m0);
                    tma_load_2d(panel0 + PANEL_BYTES, &tma_r, res_bar, n0 + 64, m0);
                }
                __syncwarp();
            }
            mbar_wait(smem_u32(&bars->tmem_full[acc]), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
            uint32_t v[2][32];
            tmem_ld_32x32(t_row, v[0]);
            if (EPI == EPI_BIAS_RESIDUAL && PANELS == 2) {
                mbar_wait(res_bar, res_phase);
                res_phase ^= 1u;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {           // 4 chunks of 32 columns = 2 panels of 64
                const uint32_t panel = panel0 + (PANELS == 2 ? (c >> 1) * PANEL_BYTES : 0);
                const uint32_t my_row = panel + lane * 128;      // this thread's accumulator row inside the panel
                if (PANELS == 1 && (c & 1) == 0) {
                    // the previous TMA store must have finished reading the panel before it is overwritten
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    if (EPI == EPI_BIAS_RESIDUAL && lane == 0) {
                        mbar_arrive_expect_tx(res_bar, PANEL_BYTES);
                        tma_load_2d(panel, &tma_r, res_bar, n0 + (c >> 1) * 64, m0);
                    }
                }
                tmem_ld_wait(v[c & 1]);
                if (c + 1 < 4) {
                    tmem_ld_32x32(t_row + (c + 1) * 32, v[(c + 1) & 1]);   // next chunk in flight during the math
                } else {
                    // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(smem_u32(&bars->tmem_empty[acc]));
                }
                if (EPI == EPI_BIAS_RESIDUAL && PANELS == 1 && (c & 1) == 0) {
                    mbar_wait(res_bar, res_phase);
                    res_phase ^= 1u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float2 f[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) f[e] = make_float2(__uint_as_float(v[c & 1][j + 2 * e]), __uint_as_float(v[c & 1][j + 2 * e + 1]));
                    if (bias != nullptr) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c * 32 + j));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c * 32 + j + 4));
                        f[0] = add_f32x2(f[0], make_float2(b0.x, b0.y));
                        f[1] = add_f32x2(f[1], make_float2(b0.z, b0.w));
                        f[2] = add_f32x2(f[2], make_float2(b1.x, b1.y));
                        f[3] = add_f32x2(f[3], make_float2(b1.z, b1.w));
                    }
                    if (EPI == EPI_BIAS_GELU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) f[e] = gelu_erf2(f[e]);
                    }
                    // 16-byte chunk k of row r sits at r*128 + ((k ^ (r & 7)) * 16) under SWIZZLE_128B
                    const uint32_t addr = my_row + ((((c & 1) * 4 + (j >> 3)) ^ (lane & 7)) << 4);
                    if (EPI == EPI_BIAS_RESIDUAL) {
                        const uint4 rr = lds128(addr);
                        f[0] = add_f32x2(f[0], unpack_bf16(rr.x));
                        f[1] = add_f32x2(f[1], unpack_bf16(rr.y));
                        f[2] = add_f32x2(f[2], unpack_bf16(rr.z));
                        f[3] = add_f32x2(f[3], unpack_bf16(rr.w));
                    }
                    uint4 o;
                    o.x = pack_bf16(f[0].x, f[0].y);
                    o.y = pack_bf16(f[1].x, f[1].y);
                    o.z = pack_bf16(f[2].x, f[2].y);
                    o.w = pack_bf16(f[3].x, f[3].y);
                    sts128(addr, o);
                }
                if (c & 1) {
                    fence_proxy_async_smem();       // generic-proxy panel writes -> visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tma_c, panel, n0 + (c >> 1) * 64, m0);
                        bulk_commit();
                    }
                }
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) bulk_wait<0>();   // all of this warp's stores have landed before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems, uint32_t box_rows) {
    static_assert(BK == 64, "the shared tensor-map helper encodes 64-element (128-byte) boxes");
    return make_tensor_map_bf16_sw128(map, base, inner, rows, row_stride_elems, box_rows);
}

template <int EPI>
int launch_t(const GemmArgs& g, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mr,
             int grid, cudaStream_t stream) {
    return launch_pdl(gemm_tc_kernel<EPI>, dim3(grid), dim3(NUM_THREADS), (size_t)smem_bytes<EPI>(), stream, ma, mb, mc, mr, g.bias, g.M, g.N, g.K);
}

}  // namespace

int gemm_tc_init() {
    if (int rc = tensormap_init()) return rc;
    cudaError_t e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<EPI_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<EPI_BIAS>());
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<EPI_BIAS_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<EPI_BIAS_GELU>());
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<EPI_BIAS_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<EPI_BIAS_RESIDUAL>());
    return (int)e;
}

int gemm_tc_launch(const GemmArgs& g, int num_sms, cudaStream_t stream) {
    if (g.M <= 0) return 0;
    if (g.N % BN != 0 || g.K % BK != 0 || (g.lda * 2) % 16 != 0 || (g.ldc % 8) != 0) return (int)cudaErrorInvalidValue;
    if (g.epilogue == EPI_BIAS_RESIDUAL && (g.R == nullptr || (g.ldr % 8) != 0)) return (int)cudaErrorInvalidValue;
    CUtensorMap ma, mb;
    int rc = make_map(&ma, g.A, (uint64_t)g.K, (uint64_t)g.a_rows_alloc, (uint64_t)g.lda, BM);
    if (rc) return rc;
    rc = make_map(&mb, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.K, BN);
    if (rc) return rc;
    CUtensorMap mc, mr;
    rc = make_map(&mc, g.C, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc, 32);
    if (rc) return rc;
    mr = mc;
    if (g.epilogue == EPI_BIAS_RESIDUAL) {
        rc = make_map(&mr, g.R, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldr, 32);
        if (rc) return rc;
    }
    const int n_tiles = ((g.M + BM - 1) / BM) * (g.N / BN);
    const int grid = n_tiles < num_sms ? n_tiles : num_sms;
    switch (g.epilogue) {
        case EPI_BIAS: return launch_t<EPI_BIAS>(g, ma, mb, mc, mr, grid, stream);
        case EPI_BIAS_GELU: return launch_t<EPI_BIAS_GELU>(g, ma, mb, mc, mr, grid, stream);
        case EPI_BIAS_RESIDUAL: return launch_t<EPI_BIAS_RESIDUAL>(g, ma, mb, mc, mr, grid, stream);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace loco
