// Shared device helpers for the loco_asr_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "loco_asr_b200 kernels are written for sm_100a (Blackwell) only"
#endif

namespace loco {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

constexpr int kHidden = 768;
constexpr int kHeads = 12;
constexpr int kHeadDim = 64;
constexpr int kFfn = 3072;
constexpr int kConvDim = 512;
constexpr int kMaxRel = 160;       // encoder_max_relative_position
constexpr int kRelCols = 320;      // rows of pe_k
constexpr int kPosK = 128;         // num_conv_pos_embeddings
constexpr int kPosGroups = 16;
constexpr int kPosGroupCh = 48;    // 768 / 16
constexpr float kLnEps = 1e-5f;

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
// GELU (ACT2FN["gelu"], exact-erf form) evaluated as  v * sigmoid(2 u(v)),  u an odd quintic fitted so that
// tanh(u(v)) == erf(v / sqrt 2) to 2.5e-5 in GELU units over the whole real line (tests/test_host_logic.py;
// v^2 is clamped at 64 where the sigmoid has long saturated, keeping u monotone).  The sigmoid form has no
// cancellation on the negative side.  9 FP32 instructions + 2 MUFU (ex2, rcp) instead of ~17 for a rational
// erf -- the GELU epilogues (conv layers, FFN1) are issue-bound, not MMA-bound, so this is what sets their speed.
// Max |error| vs exact: 2.6e-5 absolute, two orders below the bf16 rounding of the stored activation.
__device__ __forceinline__ float ex2_approx(float x) {
#ifdef LOCO_EX2_PRECISE        // tools/parity_toggles.py: what the MUFU approximation costs in parity (nothing measurable)
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
__device__ __forceinline__ float gelu_erf(float v) {
    const float v2 = fminf(v * v, 64.0f);
    float t = fmaf(v2, 0.0010142630552579922f, -0.10677572400266595f);
    t = fmaf(v2, t, -2.3011213394570755f);
    return __fdividef(v, 1.0f + ex2_approx(v * t));
}

// packed pair arithmetic (Blackwell FADD2) and the mixed-precision add (FHADD: fp32 + fp16 -> fp32 in one instruction)
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul_f32x2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
// exp2 of a pair on the FMA pipe instead of the MUFU: x = n + f with n = round(x) taken from the low mantissa bits of x + 1.5 * 2^23,
// 2^f on [-1/2, 1/2] as a cubic (relative error 7.5e-5, fifty times below the bf16 rounding of the probabilities it feeds), and n
// added into the exponent field.  Inputs are clamped at -125 (result ~2e-38 for masked / far-away keys); inputs above 127 are not
// expected (the softmax subtracts a maximum that is at most 2^8 stale).  The attention kernels send a fixed subset of the key
// columns of every block through this path: their exponential phase is bound by the 4 MUFU results per clock of the two SM
// sub-partitions an item's rows live on, while those sub-partitions' FMA pipes idle.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -125.0f);
    x.y = fmaxf(x.y, -125.0f);
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    const float2 t = add_f32x2(x, magic);
    const float2 n = add_f32x2(t, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = fma_f32x2(n, make_float2(-1.0f, -1.0f), x);
    float2 p = fma_f32x2(f, make_float2(0.055171459913253784f, 0.055171459913253784f), make_float2(0.2426108568906784f, 0.2426108568906784f));
    p = fma_f32x2(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
    p = fma_f32x2(p, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}
// MASK bit (i & 3) set: pair i of a thread's key columns takes the FMA-pipe path (0xA = every other pair, 0x8 = one in four, 0 = none).
template <int MASK>
__device__ __forceinline__ float2 ex2_pair(float2 x, int pair) {       // `pair` is a compile-time constant at every call site
    if ((MASK >> (pair & 3)) & 1) return ex2_poly2(x);
    return make_float2(ex2_approx(x.x), ex2_approx(x.y));
}
// gelu_erf on a pair with the packed FP32 pipe (FMUL2 / FFMA2 / FADD2).
// LOCO_GELU_TANH (default): the same odd quintic u(v), evaluated as 0.5 v (1 + tanh(u)) with ONE MUFU op per element
// (tanh.approx.f32, |abs err| < 5e-4 on tanh -> < 2.5e-4 |v| on the result, below the bf16 rounding of the stored
// activation) instead of ex2 + rcp: conv0's epilogue sat at the MUFU limit of 16 results / clk / SM (frontend 2.03 -> 1.71 ms
// per step; the GEMM epilogues are hidden under the next tile's MMAs and did not change).
#ifndef LOCO_GELU_TANH
#define LOCO_GELU_TANH 1
#endif
__device__ __forceinline__ float2 gelu_erf2(float2 v) {
#ifdef LOCO_GELU_ERF           // tools/parity_toggles.py: the library erff instead of the fitted forms
    return make_float2(0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f)), 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f)));
#endif
    float2 v2 = mul_f32x2(v, v);
    v2.x = fminf(v2.x, 64.0f);
    v2.y = fminf(v2.y, 64.0f);
#if LOCO_GELU_TANH
    // u = -(v t) / (2 log2 e), constants of gelu_erf scaled by -0.34657359027997264
    float2 t = fma_f32x2(v2, make_float2(-0.00035151679f, -0.00035151679f), make_float2(0.037005646f, 0.037005646f));
    t = fma_f32x2(v2, t, make_float2(0.79750788f, 0.79750788f));
    const float2 u = mul_f32x2(v, t);
    float2 th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(u.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(u.y));
    const float2 h = mul_f32x2(v, make_float2(0.5f, 0.5f));
    return fma_f32x2(h, th, h);
#else
    float2 t = fma_f32x2(v2, make_float2(0.0010142630552579922f, 0.0010142630552579922f),
                         make_float2(-0.10677572400266595f, -0.10677572400266595f));
    t = fma_f32x2(v2, t, make_float2(-2.3011213394570755f, -2.3011213394570755f));
    const float2 x = mul_f32x2(v, t);
    const float2 d = add_f32x2(make_float2(ex2_approx(x.x), ex2_approx(x.y)), make_float2(1.0f, 1.0f));
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
    return mul_f32x2(v, r);
#endif
}
__device__ __forceinline__ float add_f32_f16(float a, unsigned short h) {
    float r;
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(h), "f"(a));
    return r;
}
// Register reallocation between warpgroups (4 consecutive warps): the loader / MMA-issuer warpgroup gives registers back, the
// softmax warpgroups take them.  The kernel launches with the __launch_bounds__ allocation (65536 / threads); the pool is what
// the dec side released.
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    bf162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    bf162 h = *reinterpret_cast<bf162*>(&u);
    return __bfloat1622float2(h);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Programmatic dependent launch (opt-in, LOCO_PDL=1, see launch_pdl in internal.h; both are no-ops in a normal launch): launch_dependents lets the NEXT kernel's CTAs be scheduled as soon as SM resources free up, so
// its launch latency and prologue (barrier init, TMEM allocation, tensor-map prefetch) overlap this kernel's tail wave;
// wait blocks until the PREVIOUS kernel has completed and its memory is visible -- it must precede every global access.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe (try_wait may park the thread for a hardware-defined time; test_wait returns at once): for warps that
// poll several barriers and act on whichever completes first.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: a failed probe parks the warp (NANOSLEEP.SYNCS: until the barrier's phase completes or
// `ns` nanoseconds pass) instead of returning after ~50 cycles, so a waiting warp stops competing for issue slots with the
// warps doing the work -- it matters in the persistent kernels, whose idle roles would otherwise poll at full speed.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug traps (surfacing as a CUDA error at the next sync) instead of hanging
// the GPU box.  The bound is minutes of wall clock, far beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("loco: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, bar, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* desc, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// 2-D tiled TMA store shared -> global (bulk async-group completion); rows / columns outside the tensor are clipped.
__device__ __forceinline__ void tma_store_2d(const void* desc, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 1-D bulk copy global -> shared through the TMA engine (no tensor map), completion on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
                 : "memory");
}

// one lane of a converged warp (the compiler keeps the guarded operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[TMEM] * B[smem desc]: the A operand (bf16, K-major, two elements per 32-bit column, row m in lane m)
// is read straight from tensor memory -- how the attention kernel feeds P back into the P.V product.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all tcgen05 ops previously issued by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// N consecutive 32-bit columns of this warp's 32 lanes, N = 8 / 16 / 32 (one row per thread)
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "tmem_ld_cols");
    if constexpr (N == 32) tmem_ld_32x32(taddr, r);
    else if constexpr (N == 16) tmem_ld_32x16(taddr, r);
    else tmem_ld_32x8(taddr, r);
}
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&r)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "tmem_st_cols");
    if constexpr (N == 32) tmem_st_32x32(taddr, r);
    else if constexpr (N == 16) tmem_st_32x16(taddr, r);
    else tmem_st_32x8(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Wait for the outstanding tcgen05.ld's, then pin `r` behind the wait: the empty volatile asm statements give the
// compiler a data dependency it cannot hoist above the wait (the loads write the registers asynchronously).
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i]));
}

// Shared-memory matrix descriptor for a K-major operand tile written by TMA with SWIZZLE_128B:
// rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), version 1 (Blackwell), layout type 2.
// Bit layout: cute/arch/mma_sm100_desc.hpp `SmemDescriptor` (start [0,14), LBO [16,30), SBO [32,46),
// version [46,48), layout_type [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO
    d |= (uint64_t)1 << 46;                 // version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// MN-major operand tile written by TMA with SWIZZLE_128B: 64 MN-elements (128 B) contiguous per K row, 8-row K groups
// 1024 B apart (SBO); a single 64-wide MN atom, so LBO (stride between MN atoms) is unused.  Pairs with major bit = 1
// in the instruction descriptor.  (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>, LayoutType::B128.)
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (unused: one MN atom)
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 K-rows x 128 B
    d |= (uint64_t)1 << 46;                 // version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// Same descriptor family without swizzle (layout type 0), K-major: 8-row x 16-byte core matrices; rows of a core
// matrix 16 B apart, 8-row groups `sbo` bytes apart, the two 8-element K halves of one MMA `lbo` bytes apart
// (canonical layout ((8,m),(8,2)):((16B,SBO),(2B,LBO)), cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>).
__device__ __forceinline__ uint64_t umma_desc_noswizzle_kmajor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;                 // version
    return d;
}
// Instruction descriptor, kind::f16: bf16 A/B (format 1), fp32 accumulate, both operands K-major.
// Bit layout: `InstrDescriptor` in the same header (c_format [4,6), a_format [7,10), b_format [10,13),
// n>>3 at [17,23), m>>4 at [24,29)).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// legacy tensor path (mma.sync) used by attention and the positional conv in round 1
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, bool pred) {
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace loco
