// DEBUG ONLY: a plain shared-memory-tiled SIMT GEMM with the same contract as the tcgen05 kernel.
// It exists so the GPU tests can (a) cross-check the tensor-core path on identical inputs and (b) keep
// validating every other stage should a tcgen05 change regress.  `loco_debug_set(h, "gemm_impl", 1)`
// selects it; the product default is the tcgen05 kernel and nothing switches automatically.
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {
constexpr int TM = 64, TN = 64, TK = 16;

template <int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const bf16* __restrict__ A, int64_t lda, const bf16* __restrict__ W,
                                                        bf16* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                                        const bf16* __restrict__ R, int64_t ldr, int M, int N, int K) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sa[TK][TM + 1];
    __shared__ float sb[TK][TN + 1];
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            const int r = i / TK, c = i % TK;
            const int gm = m0 + r;
            sa[c][r] = gm < M ? __bfloat162float(A[(int64_t)gm * lda + k0 + c]) : 0.f;
            sb[c][r] = __bfloat162float(W[(int64_t)(n0 + r) * K + k0 + c]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sa[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sb[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            float v = acc[i][j];
            if (bias) v += bias[gn];
            if (EPI == EPI_BIAS_GELU) v = gelu_erf(v);
            if (EPI == EPI_BIAS_RESIDUAL) v += __bfloat162float(R[(int64_t)gm * ldr + gn]);
            C[(int64_t)gm * ldc + gn] = __float2bfloat16(v);
        }
    }
}
}  // namespace

int gemm_simt_launch(const GemmArgs& g, cudaStream_t stream) {
    if (g.M <= 0) return 0;
    if (g.N % TN != 0 || g.K % TK != 0) return (int)cudaErrorInvalidValue;
    dim3 grid((g.M + TM - 1) / TM, g.N / TN);
    switch (g.epilogue) {
        case EPI_BIAS:
            launch_pdl(gemm_simt_kernel<EPI_BIAS>, grid, dim3(256), 0, stream, g.A, g.lda, g.W, g.C, g.ldc, g.bias, g.R, g.ldr, g.M, g.N, g.K);
            break;
        case EPI_BIAS_GELU:
            launch_pdl(gemm_simt_kernel<EPI_BIAS_GELU>, grid, dim3(256), 0, stream, g.A, g.lda, g.W, g.C, g.ldc, g.bias, g.R, g.ldr, g.M, g.N, g.K);
            break;
        case EPI_BIAS_RESIDUAL:
            launch_pdl(gemm_simt_kernel<EPI_BIAS_RESIDUAL>, grid, dim3(256), 0, stream, g.A, g.lda, g.W, g.C, g.ldc, g.bias, g.R, g.ldr, g.M, g.N, g.K);
            break;
        default: return (int)cudaErrorInvalidValue;
    }
    return (int)cudaGetLastError();
}

}  // namespace loco
