// CTA-pair variant of the tcgen05 GEMM (gemm_tcgen05.cu): clusters of two CTAs on one TPC run `tcgen05.mma.cta_group::2`
// on a 256 (M) x 256 (N) x 64 (K) tile.  Each CTA loads its own 128 rows of A and HALF of the B tile (128 of the 256 weight
// rows); the pair's tensor cores read both halves, so the B operand's L2 -> shared-memory traffic and shared-memory footprint
// per CTA halve (32 KB per stage; 5 stages + two staging panels per epilogue warp), which is what the power-capped GEMMs of this encoder need:
// fewer bytes moved per FLOP.  Each CTA keeps the accumulator of its own 128 rows in its own TMEM and runs the same epilogue
// as the single-CTA kernel.
//
// Protocol (per stage s, per accumulator stage a):
//   full[s]        on the LEADER (cluster rank 0): 1 arrival (leader's expect_tx of both CTAs' bytes) + the complete_tx of all
//                  four TMA loads -- the peer's loads signal the leader's barrier (cta_group::2 TMA, peer bit cleared)
//   empty[s]       in BOTH CTAs: multicast tcgen05.commit from the leader once the MMAs that read the stage have retired
//   tmem_full[a]   in BOTH CTAs: multicast commit after the tile's last MMA
//   tmem_empty[a]  on the leader: 2 x 256 arrivals, the peer's epilogue threads arrive remotely
// The default GEMM (loco_debug_set("gemm_impl", 0) selects the single-CTA kernel); tools/gemm_sweep.py / bench.py --gemm-impl
// for the A/B numbers: in the SLURP-shaped step the GEMM time drops 14.9 -> 14.4 ms.
#include <stdio.h>

#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int BM = 128;                 // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BK = 64;
#ifndef G2_STAGES
#define G2_STAGES 5
#endif
#ifndef G2_PANELS
#define G2_PANELS 2
#endif
constexpr int STAGES = G2_STAGES;
constexpr int PANELS = G2_PANELS;         // staging panels per epilogue warp (2: the tile's two 64-column halves never wait on each other)
constexpr int A_STAGE_BYTES = BM * BK * 2;         // 16 KB
constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;   // 16 KB: this CTA's half of the weight tile
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int PANEL_BYTES = 32 * 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + PANELS * NUM_EPI_WARPS * PANEL_BYTES + 1024 + 256;
static_assert(SMEM_BYTES <= 232448, "gemm_tc2: shared memory budget");
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;        // clears the CTA-rank bit of a shared::cluster address -> the pair's leader

struct __align__(8) Barriers2 {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full[ACC_STAGES];
    uint64_t tmem_empty[ACC_STAGES];
    uint64_t res_full[NUM_EPI_WARPS];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* desc, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_result_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

struct LnArgs {                       // deferred-LayerNorm operands (GemmArgs of the same names)
    const float* stats_in;
    const float* c1;
    const float* gamma;
    float* stats_out;
};

// Tile order: N fastest -- the CTA pairs that run at the same time share A row blocks (read from DRAM once) and the whole weight
// matrix stays L2-hot.  -DG2_RASTER_N_SLOW (measurement only) walks M fastest instead: a pair keeps one weight tile and its epilogue
// vectors for consecutive tiles, but A is streamed from DRAM once per N tile (profiles/r04m_gemm_raster_experiment.txt).
#ifdef G2_RASTER_N_SLOW
#define TILE_M(t) ((t) % n_tiles_m)
#define TILE_N(t) ((t) / n_tiles_m)
#else
#define TILE_M(t) ((t) / n_tiles_n)
#define TILE_N(t) ((t) % n_tiles_n)
#endif

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r,
                const float* __restrict__ bias, const LnArgs ln, int M, int N, int K) {
    constexpr bool kRes = EPI == EPI_BIAS_RESIDUAL || EPI == EPI_BIAS_RESIDUAL_STATS || EPI == EPI_BIAS_LNRESIDUAL_STATS;
    constexpr bool kGelu = EPI == EPI_BIAS_GELU || EPI == EPI_LN_BIAS_GELU;
    constexpr bool kLnIn = EPI == EPI_LN_BIAS || EPI == EPI_LN_BIAS_GELU;          // rows of A are un-normalised: see GemmEpilogue
    constexpr bool kLnRes = EPI == EPI_BIAS_LNRESIDUAL_STATS;                      // rows of R are un-normalised
    constexpr bool kStats = EPI == EPI_BIAS_RESIDUAL_STATS || EPI == EPI_BIAS_LNRESIDUAL_STATS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
    Barriers2* bars = reinterpret_cast<Barriers2*>(smem_aligned + STAGES * STAGE_BYTES + PANELS * NUM_EPI_WARPS * PANEL_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_tiles_n = N / BN;
    const int n_tiles_m = (M + 2 * BM - 1) / (2 * BM);          // 256-row tiles
    const int n_tiles = n_tiles_m * n_tiles_n;
    const int n_kb = K / BK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_c);
        if (kRes) tma_prefetch_desc(&tma_r);
        for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(smem_u32(&bars->res_full[w]), 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(smem_u32(&bars->tmem_full[a]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[a]), 2 * NUM_EPI_WARPS * 32);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    cluster_sync_all();                  // both CTAs' barriers exist before anything is signalled across the pair
    if (warp == 2) tmem_alloc_pair(smem_u32(&bars->tmem_base), TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();                  // ... and both accumulators are allocated before the leader's first MMA
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = TILE_M(tile) * (2 * BM) + (int)rank * BM;
                const int n0 = TILE_N(tile) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1u);
                    const uint32_t full = smem_u32(&bars->full[stage]);
#ifdef G2_EXP_SKIP_B      // measurement only (wrong results): every other k-block reuses the stale weight half -> 25 % less SM ingest
                    const bool skip_b = (kb & 1) != 0;
#else
                    constexpr bool skip_b = false;
#endif
                    if (rank == 0) mbar_arrive_expect_tx(full, skip_b ? 2 * A_STAGE_BYTES : 2 * STAGE_BYTES);       // both CTAs' A tile and B half
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    tma_load_2d_pair(sa, &tma_a, full & kPeerMask, kb * BK, m0);
                    if (!skip_b) tma_load_2d_pair(sa + A_STAGE_BYTES, &tma_b, full & kPeerMask, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread of the leader CTA =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
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
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_pair(smem_u32(&bars->empty[stage]));
                    if (kb == n_kb - 1) umma_commit_pair(smem_u32(&bars->tmem_full[acc]));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs): this CTA's 128 rows of the pair's tile =====================
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const uint32_t panel0 = smem_base + STAGES * STAGE_BYTES + (warp - 4) * PANELS * PANEL_BYTES;
        const uint32_t res_bar = smem_u32(&bars->res_full[warp - 4]);
        uint32_t res_phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
            const int m0 = TILE_M(tile) * (2 * BM) + (int)rank * BM + q * 32;
            const int n0 = TILE_N(tile) * BN + half * (BN / 2);
            if (kRes && PANELS == 2) {
                // both residual panels of this warp's 32 x 128 slice, fetched while the tile's MMAs still run
                if (lane == 0) {
                    bulk_wait_read<0>();            // the previous tile's stores have finished reading the panels
                    mbar_arrive_expect_tx(res_bar, 2 * PANEL_BYTES);
                    tma_load_2d(panel0, &tma_r, res_bar, n0, m0);
                    tma_load_2d(panel0 + PANEL_BYTES, &tma_r, res_bar, n0 + 64, m0);
                }
                __syncwarp();
            }
            // deferred LayerNorm: this thread's row as x * ra + rb = (x - mean) * rstd, from the six (mean, M2) partials
            float2 ra2 = make_float2(1.f, 1.f), rb2 = make_float2(0.f, 0.f);
            if (kLnIn || kLnRes) {
                const int row = min(m0 + lane, M - 1);
                const float4* sp = reinterpret_cast<const float4*>(ln.stats_in + (int64_t)row * (2 * kStatSlots));
                const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);
                const float mean = (s0.x + s0.z + s1.x + s1.z + s2.x + s2.z) * (1.0f / kStatSlots);
                float m2 = (s0.y + s0.w) + (s1.y + s1.w) + (s2.y + s2.w);
                const float d0 = s0.x - mean, d1 = s0.z - mean, d2 = s1.x - mean, d3 = s1.z - mean, d4 = s2.x - mean, d5 = s2.z - mean;
                m2 = fmaf(128.f, fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, fmaf(d4, d4, d5 * d5))))), m2);
                const float rstd = rsqrtf(m2 * (1.0f / 768.f) + kLnEps);
                ra2 = make_float2(rstd, rstd);
                rb2 = make_float2(-mean * rstd, -mean * rstd);
            }
            float2 ssum = make_float2(0.f, 0.f), ssq = make_float2(0.f, 0.f);
            mbar_wait(smem_u32(&bars->tmem_full[acc]), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
            uint32_t v[2][32];
            tmem_ld_32x32(t_row, v[0]);
            if (kRes && PANELS == 2) {
                mbar_wait(res_bar, res_phase);
                res_phase ^= 1u;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t panel = panel0 + (PANELS == 2 ? (c >> 1) * PANEL_BYTES : 0);
                const uint32_t my_row = panel + lane * 128;
                if ((c & 1) == 0 && (PANELS == 1 || !kRes)) {
                    // the TMA store that last read this panel must have finished before it is overwritten
                    if (lane == 0) {
                        if (PANELS == 2) bulk_wait_read<1>();      // only the OTHER panel's store (the most recent group) may be in flight
                        else bulk_wait_read<0>();
                    }
                    __syncwarp();
                    if (kRes && lane == 0) {
                        mbar_arrive_expect_tx(res_bar, PANEL_BYTES);
                        tma_load_2d(panel, &tma_r, res_bar, n0 + (c >> 1) * 64, m0);
                    }
                }
                tmem_ld_wait(v[c & 1]);
                if (c + 1 < 4) {
                    tmem_ld_32x32(t_row + (c + 1) * 32, v[(c + 1) & 1]);
                } else {
                    tc_fence_before();
                    mbar_arrive_cluster(smem_u32(&bars->tmem_empty[acc]) & kPeerMask);      // the leader's barrier
                }
                if (kRes && PANELS == 1 && (c & 1) == 0) {
                    mbar_wait(res_bar, res_phase);
                    res_phase ^= 1u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float2 f[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) f[e] = make_float2(__uint_as_float(v[c & 1][j + 2 * e]), __uint_as_float(v[c & 1][j + 2 * e + 1]));
                    if (kLnIn) {
                        // rstd_r (acc - mean_r c1[n]) + bias[n]
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c * 32 + j));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c * 32 + j + 4));
                        const float4 k0 = __ldg(reinterpret_cast<const float4*>(ln.c1 + n0 + c * 32 + j));
                        const float4 k1 = __ldg(reinterpret_cast<const float4*>(ln.c1 + n0 + c * 32 + j + 4));
                        f[0] = fma_f32x2(f[0], ra2, fma_f32x2(make_float2(k0.x, k0.y), rb2, make_float2(b0.x, b0.y)));
                        f[1] = fma_f32x2(f[1], ra2, fma_f32x2(make_float2(k0.z, k0.w), rb2, make_float2(b0.z, b0.w)));
                        f[2] = fma_f32x2(f[2], ra2, fma_f32x2(make_float2(k1.x, k1.y), rb2, make_float2(b1.x, b1.y)));
                        f[3] = fma_f32x2(f[3], ra2, fma_f32x2(make_float2(k1.z, k1.w), rb2, make_float2(b1.z, b1.w)));
                    } else if (bias != nullptr) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c * 32 + j));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c * 32 + j + 4));
                        f[0] = add_f32x2(f[0], make_float2(b0.x, b0.y));
                        f[1] = add_f32x2(f[1], make_float2(b0.z, b0.w));
                        f[2] = add_f32x2(f[2], make_float2(b1.x, b1.y));
                        f[3] = add_f32x2(f[3], make_float2(b1.z, b1.w));
                    }
                    if (kGelu) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) f[e] = gelu_erf2(f[e]);
                    }
                    const uint32_t addr = my_row + ((((c & 1) * 4 + (j >> 3)) ^ (lane & 7)) << 4);
                    if (kRes) {
                        const uint4 rr = lds128(addr);
                        float2 r[4] = {unpack_bf16(rr.x), unpack_bf16(rr.y), unpack_bf16(rr.z), unpack_bf16(rr.w)};
                        if (kLnRes) {
                            // the residual is LayerNorm(R) = ((R - mean_r) rstd_r) gamma[n] + beta[n]; beta travels inside `bias`
                            const float4 g0 = __ldg(reinterpret_cast<const float4*>(ln.gamma + n0 + c * 32 + j));
                            const float4 g1 = __ldg(reinterpret_cast<const float4*>(ln.gamma + n0 + c * 32 + j + 4));
                            f[0] = fma_f32x2(fma_f32x2(r[0], ra2, rb2), make_float2(g0.x, g0.y), f[0]);
                            f[1] = fma_f32x2(fma_f32x2(r[1], ra2, rb2), make_float2(g0.z, g0.w), f[1]);
                            f[2] = fma_f32x2(fma_f32x2(r[2], ra2, rb2), make_float2(g1.x, g1.y), f[2]);
                            f[3] = fma_f32x2(fma_f32x2(r[3], ra2, rb2), make_float2(g1.z, g1.w), f[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) f[e] = add_f32x2(f[e], r[e]);
                        }
                    }
                    if (kStats) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            ssum = add_f32x2(ssum, f[e]);
                            ssq = fma_f32x2(f[e], f[e], ssq);
                        }
                    }
                    uint4 o;
                    o.x = pack_bf16(f[0].x, f[0].y);
                    o.y = pack_bf16(f[1].x, f[1].y);
                    o.z = pack_bf16(f[2].x, f[2].y);
                    o.w = pack_bf16(f[3].x, f[3].y);
                    sts128(addr, o);
                }
                if (c & 1) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tma_c, panel, n0 + (c >> 1) * 64, m0);
                        bulk_commit();
                    }
                }
            }
            if (kStats) {
                // (mean, M2) of this row's 128 columns; slot = position of the slice in the 768-wide row
                const float sum = ssum.x + ssum.y, sq = ssq.x + ssq.y;
                const float mean_p = sum * (1.0f / 128.f);
                const int row = m0 + lane;
                if (row < M)
                    *reinterpret_cast<float2*>(ln.stats_out + (int64_t)row * (2 * kStatSlots) + 2 * (TILE_N(tile) * 2 + half)) =
                        make_float2(mean_p, fmaxf(sq - sum * mean_p, 0.f));
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) bulk_wait<0>();
    }

    tc_fence_before();
    cluster_sync_all();                  // both CTAs have finished with both accumulators
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

template <int EPI>
int launch2_t(const GemmArgs& g, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mr, int grid,
              cudaStream_t stream) {
    const LnArgs ln = {g.stats_in, g.c1, g.ln_gamma, g.stats_out};
    return launch_pdl(gemm_tc2_kernel<EPI>, dim3(grid), dim3(NUM_THREADS), (size_t)SMEM_BYTES, stream, ma, mb, mc, mr, g.bias, ln, g.M, g.N, g.K);
}

}  // namespace

int gemm_tc2_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI_BIAS_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI_BIAS_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI_LN_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI_LN_BIAS_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI_BIAS_RESIDUAL_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI_BIAS_LNRESIDUAL_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    return (int)e;
}

int gemm_tc2_launch(const GemmArgs& g, int num_sms, cudaStream_t stream) {
    if (g.M <= 0) return 0;
    if (g.N % BN != 0 || g.K % BK != 0 || (g.lda * 2) % 16 != 0 || (g.ldc % 8) != 0) return (int)cudaErrorInvalidValue;
    const bool has_res = g.epilogue == EPI_BIAS_RESIDUAL || g.epilogue == EPI_BIAS_RESIDUAL_STATS || g.epilogue == EPI_BIAS_LNRESIDUAL_STATS;
    const bool ln_in = g.epilogue == EPI_LN_BIAS || g.epilogue == EPI_LN_BIAS_GELU;
    const bool stats = g.epilogue == EPI_BIAS_RESIDUAL_STATS || g.epilogue == EPI_BIAS_LNRESIDUAL_STATS;
    if (has_res && (g.R == nullptr || (g.ldr % 8) != 0)) return (int)cudaErrorInvalidValue;
    if (ln_in && (!g.stats_in || !g.c1 || !g.bias)) return (int)cudaErrorInvalidValue;
    if (stats && (!g.stats_out || g.N != kStatSlots * 128)) return (int)cudaErrorInvalidValue;
    if (g.epilogue == EPI_BIAS_LNRESIDUAL_STATS && (!g.stats_in || !g.ln_gamma || !g.bias)) return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap ma, mb, mc, mr;
    int rc = make_tensor_map_bf16_sw128(&ma, g.A, (uint64_t)g.K, (uint64_t)g.a_rows_alloc, (uint64_t)g.lda, BM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_sw128(&mb, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.K, BN / 2);
    if (rc) return rc;
    rc = make_tensor_map_bf16_sw128(&mc, g.C, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc, 32);
    if (rc) return rc;
    mr = mc;
    if (has_res) {
        rc = make_tensor_map_bf16_sw128(&mr, g.R, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldr, 32);
        if (rc) return rc;
    }
    const int n_tiles = ((g.M + 2 * BM - 1) / (2 * BM)) * (g.N / BN);
    int grid = 2 * (n_tiles < num_sms / 2 ? n_tiles : num_sms / 2);
    switch (g.epilogue) {
        case EPI_BIAS: return launch2_t<EPI_BIAS>(g, ma, mb, mc, mr, grid, stream);
        case EPI_BIAS_GELU: return launch2_t<EPI_BIAS_GELU>(g, ma, mb, mc, mr, grid, stream);
        case EPI_BIAS_RESIDUAL: return launch2_t<EPI_BIAS_RESIDUAL>(g, ma, mb, mc, mr, grid, stream);
        case EPI_LN_BIAS: return launch2_t<EPI_LN_BIAS>(g, ma, mb, mc, mr, grid, stream);
        case EPI_LN_BIAS_GELU: return launch2_t<EPI_LN_BIAS_GELU>(g, ma, mb, mc, mr, grid, stream);
        case EPI_BIAS_RESIDUAL_STATS: return launch2_t<EPI_BIAS_RESIDUAL_STATS>(g, ma, mb, mc, mr, grid, stream);
        case EPI_BIAS_LNRESIDUAL_STATS: return launch2_t<EPI_BIAS_LNRESIDUAL_STATS>(g, ma, mb, mc, mr, grid, stream);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace loco
