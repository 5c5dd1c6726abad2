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

void decode_mlp(const float* shifts, const float* code, int64_t V, const float* w1, const float* b1, const float* w2,
                const float* b2, const float* w3, const float* signed_scale, float* values, float* grad,
                cudaStream_t s) {
    if (V == 0) return;
    const unsigned blocks = (unsigned)std::min<size_t>(grid_for((size_t)V * 32, 256), 148 * 8);
    ProfileScope prof("decode_mlp", s, (double)V * 2.0 * (35 * 32 + 32 * 32 + 64));
    decode_kernel<<<blocks, 256, 0, s>>>(shifts, code, V, w1, b1, w2, b2, w3, signed_scale, values, grad);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
