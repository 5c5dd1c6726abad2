// D[M,N] = A[M,K] @ W[K,N] in fp32 accuracy on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, 3xTF32 split, accumulators in TMEM).
//
// Used for the dense decoder MLP between the U-Net and contouring (reference
// UNet5.decode, net_definitions_torch.py:655-666) and as the validated building
// block of the tensor-core sparse convolution.  One CTA = 128 rows x N columns
// (N <= 256, multiple of 16); K is consumed in chunks of 32.  W is pre-packed
// once into the K-major no-swizzle canonical layout (hi and lo parts) so a chunk
// of B is one contiguous copy; A rows are split into hi/lo while being staged.
#include "internal.h"
#include "profile.cuh"
#include "umma.cuh"

namespace asrb {

// packed weights: [chunks][2 (hi, lo)][N rows x 32 k] tiles in canonical layout
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ W, int K, int N, int n_pad, float* __restrict__ out) {
    const int chunks = (K + 31) / 32;
    const long long total = (long long)chunks * n_pad * 32;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / (n_pad * 32));
        const int rem = (int)(i % (n_pad * 32));
        const int n = rem / 32, kk = rem % 32;
        const int k = c * 32 + kk;
        const float w = (k < K && n < N) ? W[(size_t)k * N + n] : 0.f;
        const float hi = umma::tf32_hi(w);
        const size_t tile = (size_t)c * 2 * n_pad * 32;
        const size_t off = ((size_t)(n >> 3) * 256) + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3);
        out[tile + off] = hi;
        out[tile + (size_t)n_pad * 32 + off] = w - hi;
    }
}

template <int EPI>  // 0: plain store, 1: + bias, ReLU
__global__ void __launch_bounds__(128)
dense_gemm_kernel(const float* __restrict__ A, long long M, int K, int lda, const float* __restrict__ Wp, int n_pad,
                  int N, const float* __restrict__ bias, float* __restrict__ D, int ldd) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = smem + 128 * 128;
    uint8_t* sB = smem + 2 * 128 * 128;  // hi then lo, n_pad*128 bytes each
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long row0 = blockIdx.x * 128LL;
    const uint32_t ncols = n_pad <= 32 ? 32 : n_pad <= 64 ? 64 : n_pad <= 128 ? 128 : 256;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        umma::mbar_init(&mbar, 1);
        umma::fence_barrier_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma::make_idesc_tf32(128, n_pad);
    const int chunks = (K + 31) / 32;
    const long long r = row0 + tid;
    uint32_t phase = 0;
    for (int c = 0; c < chunks; ++c) {
        // A: thread = row; 8 float4 -> hi / lo tiles
#pragma unroll
        for (int kq = 0; kq < 8; ++kq) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = c * 32 + kq * 4 + j;
                v[j] = (r < M && k < K) ? A[(size_t)r * lda + k] : 0.f;
            }
            float4 hi, lo;
            hi.x = umma::tf32_hi(v[0]); hi.y = umma::tf32_hi(v[1]); hi.z = umma::tf32_hi(v[2]); hi.w = umma::tf32_hi(v[3]);
            lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
            const uint32_t off = umma::tile_offset(tid, kq * 4);
            *reinterpret_cast<float4*>(sA_hi + off) = hi;
            *reinterpret_cast<float4*>(sA_lo + off) = lo;
        }
        // B: contiguous copy of the pre-packed hi|lo tiles of this chunk
        const float4* src = reinterpret_cast<const float4*>(Wp + (size_t)c * 2 * n_pad * 32);
        float4* dst = reinterpret_cast<float4*>(sB);
        for (int i = tid; i < 2 * n_pad * 8; i += 128) dst[i] = src[i];
        umma::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            umma::tc_fence_after();
            const uint32_t a_hi = umma::smem_u32(sA_hi), a_lo = umma::smem_u32(sA_lo);
            const uint32_t b_hi = umma::smem_u32(sB), b_lo = b_hi + n_pad * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t o = ks * 2 * umma::kLBO;
                umma::mma_tf32(tmem, umma::make_smem_desc(a_hi + o), umma::make_smem_desc(b_hi + o), idesc, c > 0 || ks > 0);
                umma::mma_tf32(tmem, umma::make_smem_desc(a_lo + o), umma::make_smem_desc(b_hi + o), idesc, true);
                umma::mma_tf32(tmem, umma::make_smem_desc(a_hi + o), umma::make_smem_desc(b_lo + o), idesc, true);
            }
            umma::mma_commit(&mbar);
        }
        umma::mbar_wait(&mbar, phase);  // operands consumed, accumulator up to date
        phase ^= 1;
    }
    umma::tc_fence_after();
    // epilogue: thread = row (TMEM lane), 32 columns at a time
    for (int n0 = 0; n0 < n_pad; n0 += 32) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
        if (r < M) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + j;
                if (n < N) {
                    float x = v[j];
                    if (EPI == 1) x = fmaxf(x + bias[n], 0.f);
                    D[(size_t)r * ldd + n] = x;
                }
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

size_t packed_weights_floats(int K, int N) {
    const int n_pad = ((N + 15) / 16) * 16;
    return (size_t)((K + 31) / 32) * 2 * n_pad * 32;
}

void pack_weights(const float* W, int K, int N, float* out, cudaStream_t s) {
    const int n_pad = ((N + 15) / 16) * 16;
    const long long total = (long long)((K + 31) / 32) * n_pad * 32;
    pack_weights_kernel<<<grid_for(total, 256), 256, 0, s>>>(W, K, N, n_pad, out);
    ASRB_CHECK_LAUNCH();
}

void dense_gemm_tf32x3(const float* A, int64_t M, int K, int lda, const float* Wp, int N, const float* bias, int relu,
                       float* D, int ldd, cudaStream_t s) {
    ASRB_REQUIRE(N >= 1 && N <= 256, "dense_gemm: N must be in [1, 256]");
    if (M == 0) return;
    const int n_pad = ((N + 15) / 16) * 16;
    const size_t smem = 2 * 128 * 128 + 2 * (size_t)n_pad * 128;
    ProfileScope prof("dense_gemm_tf32x3", s, 2.0 * (double)M * K * N);
    if (bias && relu) {
        ASRB_CUDA(cudaFuncSetAttribute(dense_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_gemm_kernel<1><<<grid_for(M, 128), 128, smem, s>>>(A, M, K, lda, Wp, n_pad, N, bias, D, ldd);
    } else {
        ASRB_CUDA(cudaFuncSetAttribute(dense_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_gemm_kernel<0><<<grid_for(M, 128), 128, smem, s>>>(A, M, K, lda, Wp, n_pad, N, bias, D, ldd);
    }
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
