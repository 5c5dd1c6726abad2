// Decoder MLP 35 -> 32 -> 32 -> 2 on cat(shift, code) with ReLU between layers
// (reference UNet5.decode / decode_with_gradient,
// models/v0/net_definitions_torch.py:655-686).
//
// Warp-per-voxel, lane = hidden neuron: every lane keeps its own rows of W1 and
// W2 in registers for the whole (persistent) kernel, the 35 inputs and the 32
// hidden activations are broadcast with warp shuffles, and the two outputs are
// shuffle reductions.  The code row is one coalesced 128-byte load.  The
// optional analytic gradient of the signed channel w.r.t. the shift follows the
// reference's ReLU-masked back-substitution.  `signed_scale` fuses the
// per-voxel rescaling of channel 0 that the pipeline applies before contouring
// (cpp/lib/asr.cpp:334-336).
#include "internal.h"
#include "profile.cuh"

namespace asrb {

constexpr int kIn = 35, kH = 32;

__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ shifts, const float* __restrict__ code, long long V,
              const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
              const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ signed_scale,
              float* __restrict__ values, float* __restrict__ grad) {
    __shared__ float s_w1[kH * kIn], s_w2[kH * kH];
    const int lane = threadIdx.x & 31;
    if (grad) {
        for (int i = threadIdx.x; i < kH * kIn; i += blockDim.x) s_w1[i] = w1[i];
        for (int i = threadIdx.x; i < kH * kH; i += blockDim.x) s_w2[i] = w2[i];
        __syncthreads();
    }
    float r1[kIn], r2[kH];
#pragma unroll
    for (int i = 0; i < kIn; ++i) r1[i] = w1[lane * kIn + i];
#pragma unroll
    for (int i = 0; i < kH; ++i) r2[i] = w2[lane * kH + i];
    const float bias1 = b1[lane], bias2 = b2[lane];
    const float w3s = w3[lane], w3u = w3[kH + lane];

    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long v = warp0; v < V; v += nwarps) {
        const float c = __ldg(code + (size_t)v * kH + lane);
        const float sh = (shifts && lane < 3) ? shifts[3 * v + lane] : 0.f;
        float acc = bias1;
#pragma unroll
        for (int i = 0; i < 3; ++i) acc = fmaf(__shfl_sync(0xffffffffu, sh, i), r1[i], acc);
#pragma unroll
        for (int i = 0; i < kH; ++i) acc = fmaf(__shfl_sync(0xffffffffu, c, i), r1[3 + i], acc);
        const float h1 = fmaxf(acc, 0.f);
        acc = bias2;
#pragma unroll
        for (int i = 0; i < kH; ++i) acc = fmaf(__shfl_sync(0xffffffffu, h1, i), r2[i], acc);
        const float h2 = fmaxf(acc, 0.f);
        float os = h2 * w3s, ou = h2 * w3u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            os += __shfl_xor_sync(0xffffffffu, os, d);
            ou += __shfl_xor_sync(0xffffffffu, ou, d);
        }
        if (lane == 0) {
            if (signed_scale) os *= signed_scale[v];
            reinterpret_cast<float2*>(values)[v] = make_float2(os, ou);
        }
        if (grad) {
            // d value[0] / d shift: z3 = W3[0] masked by h2 > 0; z2 = z3 W2 masked by h1 > 0; z1 = z2 W1[:, :3]
            const float z3 = h2 <= 0.f ? 0.f : w3s;
            float z2 = 0.f;
            for (int o = 0; o < kH; ++o) z2 = fmaf(__shfl_sync(0xffffffffu, z3, o), s_w2[o * kH + lane], z2);
            if (h1 <= 0.f) z2 = 0.f;
            float z1 = 0.f;
            for (int o = 0; o < kH; ++o) {
                const float z = __shfl_sync(0xffffffffu, z2, o);
                if (lane < 3) z1 = fmaf(z, s_w1[o * kIn + lane], z1);
            }
            if (lane < 3) grad[3 * v + lane] = z1;
        }
    }
}

// Value-only fast path: one THREAD per voxel.  The weights sit in shared memory and every
// access is a warp-wide broadcast (all lanes read the same weight), the activations stay in
// registers, so the kernel is a plain FMA stream (2.2 k FMAs per voxel) instead of the
// shuffle-bound warp-per-voxel form above (which remains for the gradient variant).
__global__ void __launch_bounds__(128)
decode_thread_kernel(const float* __restrict__ shifts, const float* __restrict__ code, long long V,
                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                     const float* __restrict__ b2, const float* __restrict__ w3,
                     const float* __restrict__ signed_scale, float* __restrict__ values) {
    // w1 rows padded to 36 so that a row starts 16-byte aligned: [h][36], inputs 0..2 = shift, 3..34 = code
    __shared__ __align__(16) float s_w1[kH * 36];
    __shared__ __align__(16) float s_w2[kH * kH];
    __shared__ float s_b1[kH], s_b2[kH], s_w3[2 * kH];
    for (int i = threadIdx.x; i < kH * 36; i += blockDim.x) {
        const int h = i / 36, k = i % 36;
        s_w1[i] = k < kIn ? w1[h * kIn + k] : 0.f;
    }
    for (int i = threadIdx.x; i < kH * kH; i += blockDim.x) s_w2[i] = w2[i];
    if (threadIdx.x < kH) {
        s_b1[threadIdx.x] = b1[threadIdx.x];
        s_b2[threadIdx.x] = b2[threadIdx.x];
    }
    if (threadIdx.x < 2 * kH) s_w3[threadIdx.x] = w3[threadIdx.x];
    __syncthreads();
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= V) return;
    float in[36];
    in[0] = shifts ? shifts[3 * v] : 0.f;
    in[1] = shifts ? shifts[3 * v + 1] : 0.f;
    in[2] = shifts ? shifts[3 * v + 2] : 0.f;
    in[35] = 0.f;
    const float4* crow = reinterpret_cast<const float4*>(code + (size_t)v * kH);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 c = __ldg(crow + j);
        in[3 + 4 * j] = c.x;
        in[4 + 4 * j] = c.y;
        in[5 + 4 * j] = c.z;
        in[6 + 4 * j] = c.w;
    }
    float h1[kH];
#pragma unroll
    for (int h = 0; h < kH; ++h) {
        float acc = s_b1[h];
        // same accumulation order as the reference layer: shift inputs first, then the code
#pragma unroll
        for (int k = 0; k < 36; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(s_w1 + h * 36 + k);
            acc = fmaf(in[k], w.x, acc);
            acc = fmaf(in[k + 1], w.y, acc);
            acc = fmaf(in[k + 2], w.z, acc);
            acc = fmaf(in[k + 3], w.w, acc);
        }
        h1[h] = fmaxf(acc, 0.f);
    }
    float os = 0.f, ou = 0.f;
#pragma unroll
    for (int h = 0; h < kH; ++h) {
        float acc = s_b2[h];
#pragma unroll
        for (int k = 0; k < kH; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(s_w2 + h * kH + k);
            acc = fmaf(h1[k], w.x, acc);
            acc = fmaf(h1[k + 1], w.y, acc);
            acc = fmaf(h1[k + 2], w.z, acc);
            acc = fmaf(h1[k + 3], w.w, acc);
        }
        const float h2 = fmaxf(acc, 0.f);
        os = fmaf(h2, s_w3[h], os);
        ou = fmaf(h2, s_w3[kH + h], ou);
    }
    if (signed_scale) os *= signed_scale[v];
    reinterpret_cast<float2*>(values)[v] = make_float2(os, ou);
}

void decode_mlp(const float* shifts, const float* code, int64_t V, const float* w1, const float* b1, const float* w2,
                const float* b2, const float* w3, const float* signed_scale, float* values, float* grad,
                cudaStream_t s) {
    if (V == 0) return;
    const unsigned blocks = (unsigned)std::min<size_t>(grid_for((size_t)V * 32, 256), 148 * 8);
    ProfileScope prof("decode_mlp", s, (double)V * 2.0 * (35 * 32 + 32 * 32 + 64));
    if (!grad && ((uintptr_t)code % 16) == 0) {
        decode_thread_kernel<<<grid_for((size_t)V, 128), 128, 0, s>>>(shifts, code, V, w1, b1, w2, b2, w3, signed_scale,
                                                                      values);
        ASRB_CHECK_LAUNCH();
        return;
    }
    decode_kernel<<<blocks, 256, 0, s>>>(shifts, code, V, w1, b1, w2, b2, w3, signed_scale, values, grad);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
