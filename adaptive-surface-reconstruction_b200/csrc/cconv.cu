// Continuous convolution (points -> voxel features), the aggregation step.
//
// Replaces Open3D-ML's `continuous_conv` op as the reference uses it through
// ml3d.layers.ContinuousConv (models/v0/net_definitions_torch.py:53-70,108-116):
// coordinate_mapping='ball_to_cube_radial', align_corners=True,
// interpolation='linear', normalize=True, user-supplied neighbour lists with a
// per-pair importance.  Open3D builds a dense [S^3*Cin x 32-voxel] matrix per
// block and calls an Eigen GEMM; here one warp owns one output voxel, keeps the
// S^3 x Cin cell tensor in shared memory, splats each neighbour with its 8
// trilinear taps (lane = tap x channel, so a pair is one conflict-free
// shared-memory update), and contracts with the filter in registers
// (lane = output channel).  Bias and ReLU of the layer are fused.
#include "internal.h"
#include "profile.cuh"

namespace asrb {

constexpr int kCcWarps = 8;

__global__ void __launch_bounds__(kCcWarps * 32)
cconv_kernel(const float* __restrict__ filters,  // [S,S,S,Cin,Cout]  (z,y,x order)
             const float* __restrict__ out_pos, const float* __restrict__ extents, int extents_stride,
             const float* __restrict__ offset, const float* __restrict__ inp_pos,
             const float* __restrict__ inp_feat, const float* __restrict__ inp_importance,
             const int32_t* __restrict__ nidx, const float* __restrict__ nimp,
             const int64_t* __restrict__ splits, long long V, int S, int Cin, int Cout, int normalize,
             const float* __restrict__ bias, int relu, float* __restrict__ out) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = S * S * S;
    float* B = smem + (size_t)warp * cells * Cin;
    const long long v = blockIdx.x * (long long)kCcWarps + warp;
    if (v >= V) return;
    for (int j = lane; j < cells * Cin; j += 32) B[j] = 0.f;
    __syncwarp();
    const float cx = out_pos[3 * v], cy = out_pos[3 * v + 1], cz = out_pos[3 * v + 2];
    const float scale = 2.0f / extents[v * extents_stride];
    const float ox = offset ? offset[0] : 0.f, oy = offset ? offset[1] : 0.f, oz = offset ? offset[2] : 0.f;
    const float sm1 = (float)(S - 1);
    float norm = 0.f;
    const int64_t e = splits[v + 1];
    for (int64_t n = splits[v]; n < e; ++n) {
        const int p = nidx[n];
        float imp = nimp ? nimp[n] : 1.0f;
        norm += imp;
        if (inp_importance) imp *= inp_importance[p];
        // relative position -> unit ball -> cube [-0.5, 0.5]^3 -> kernel coordinates
        float x = (inp_pos[3 * (size_t)p] - cx) * scale;
        float y = (inp_pos[3 * (size_t)p + 1] - cy) * scale;
        float z = (inp_pos[3 * (size_t)p + 2] - cz) * scale;
        const float nrm = sqrtf(x * x + y * y + z * z);
        const float amax = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        const float stretch = amax < 1e-8f ? 0.f : 0.5f * nrm / amax;
        x = fminf(fmaxf((x * stretch + ox + 0.5f) * sm1, 0.f), sm1);
        y = fminf(fmaxf((y * stretch + oy + 0.5f) * sm1, 0.f), sm1);
        z = fminf(fmaxf((z * stretch + oz + 0.5f) * sm1, 0.f), sm1);
        const int x0 = min((int)x, S - 1), y0 = min((int)y, S - 1), z0 = min((int)z, S - 1);
        const float ax = x - (float)x0, ay = y - (float)y0, az = z - (float)z0;
        for (int it = lane; it < 8 * Cin; it += 32) {
            const int tap = it / Cin, ch = it - tap * Cin;
            const int tx = tap & 1, ty = (tap >> 1) & 1, tz = tap >> 2;
            // a "+1" tap that is clamped onto its "+0" twin carries weight 0: skip
            if ((tx && x0 == S - 1) || (ty && y0 == S - 1) || (tz && z0 == S - 1)) continue;
            const float w = (tx ? ax : 1.f - ax) * (ty ? ay : 1.f - ay) * (tz ? az : 1.f - az);
            const int cell = ((z0 + tz) * S + (y0 + ty)) * S + (x0 + tx);
            B[cell * Cin + ch] += w * (inp_feat[(size_t)p * Cin + ch] * imp);
        }
        __syncwarp();
    }
    const int K = cells * Cin;
    for (int oc = lane; oc < Cout; oc += 32) {
        float acc = 0.f;
        for (int j = 0; j < K; ++j) acc = fmaf(B[j], __ldg(filters + (size_t)j * Cout + oc), acc);
        if (normalize && norm != 0.f) acc /= norm;
        if (bias) acc += bias[oc];
        if (relu) acc = fmaxf(acc, 0.f);
        out[(size_t)v * Cout + oc] = acc;
    }
}

// Fast path for the reference configuration (Cin = 4): persistent blocks keep the
// filter bank in shared memory; a warp owns a voxel and takes its neighbours in
// batches of 32 — phase A: lane = neighbour (index, position, 4 features and
// importance loaded with 32-way memory parallelism, kernel coordinates computed
// once per neighbour); phase B: the batch is replayed through warp shuffles with
// lane = tap x channel, one conflict-free shared-memory update per neighbour.
// The contraction with the filter bank runs for kCcGroup voxels of the warp at once
// (lane = output channel): every filter value read from shared memory feeds kCcGroup
// FMAs instead of one, which took the shared-memory pipe off the critical path.
// Voxels without neighbours skip the contraction (output = act(bias)).
constexpr int kCcGroup = 4;
constexpr int kCc4Warps = 16;  // 32 KB filter bank + 16 x 4 KB cell tensors = 96 KB -> 2 blocks / SM, 32 warps

template <int S>
__global__ void __launch_bounds__(kCc4Warps * 32)
cconv4_kernel(const float* __restrict__ filters, const float* __restrict__ out_pos,
              const float* __restrict__ extents, int extents_stride, const float* __restrict__ offset,
              const float* __restrict__ inp_pos, const float4* __restrict__ inp_feat,
              const float* __restrict__ inp_importance, const int32_t* __restrict__ nidx,
              const float* __restrict__ nimp, const int64_t* __restrict__ splits, long long V, int Cout,
              int normalize, const float* __restrict__ bias, int relu, float* __restrict__ out) {
    constexpr int CELLS = S * S * S, K = CELLS * 4;
    extern __shared__ float smem[];
    float* Ws = smem;                                   // [K][Cout]
    float* Bs = smem + (size_t)K * Cout;                // [warps][kCcGroup][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) Ws[i] = filters[i];
    __syncthreads();
    float* Bw = Bs + (size_t)warp * kCcGroup * K;
    float* stage = Bs + (size_t)kCc4Warps * kCcGroup * K + warp * 256;  // [32 neighbours][8]
    const float ox = offset ? offset[0] : 0.f, oy = offset ? offset[1] : 0.f, oz = offset ? offset[2] : 0.f;
    const float sm1 = (float)(S - 1);
    const int tap = lane >> 2, ch = lane & 3;
    const int tx = tap & 1, ty = (tap >> 1) & 1, tz = tap >> 2;
    const int tapmask = (tx << 8) | (ty << 9) | (tz << 10);  // flags that make this lane's tap a duplicate
    const int tapoff = (tz * S + ty) * S + tx;               // cell offset of this lane's tap
    const long long nwarps = (long long)gridDim.x * kCc4Warps;
    long long gv[kCcGroup];
    float gnorm[kCcGroup];
    int g = 0;
    auto flush = [&]() {
        // out[gv[i], :] = act(B_i . W / norm_i + bias) for the g collected voxels
        for (int oc = lane; oc < Cout; oc += 32) {
            float acc[kCcGroup];
#pragma unroll
            for (int i = 0; i < kCcGroup; ++i) acc[i] = 0.f;
#pragma unroll 2
            for (int j = 0; j < K; j += 4) {
                const float w0 = Ws[(j + 0) * Cout + oc], w1 = Ws[(j + 1) * Cout + oc], w2 = Ws[(j + 2) * Cout + oc],
                            w3 = Ws[(j + 3) * Cout + oc];
#pragma unroll
                for (int i = 0; i < kCcGroup; ++i) {
                    if (i < g) {
                        const float4 bv = *reinterpret_cast<const float4*>(Bw + i * K + j);
                        acc[i] = fmaf(bv.x, w0, acc[i]);
                        acc[i] = fmaf(bv.y, w1, acc[i]);
                        acc[i] = fmaf(bv.z, w2, acc[i]);
                        acc[i] = fmaf(bv.w, w3, acc[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kCcGroup; ++i) {
                if (i < g) {
                    float r = acc[i];
                    if (normalize && gnorm[i] != 0.f) r /= gnorm[i];
                    if (bias) r += bias[oc];
                    if (relu) r = fmaxf(r, 0.f);
                    out[(size_t)gv[i] * Cout + oc] = r;
                }
            }
        }
        __syncwarp();
        g = 0;
    };
    for (long long v = blockIdx.x * (long long)kCc4Warps + warp; v < V; v += nwarps) {
        const int64_t b = splits[v], e = splits[v + 1];
        if (e <= b) {
            for (int oc = lane; oc < Cout; oc += 32) {
                float r = bias ? bias[oc] : 0.f;
                if (relu) r = fmaxf(r, 0.f);
                out[(size_t)v * Cout + oc] = r;
            }
            continue;
        }
        float* B = Bw + g * K;
        for (int j = lane; j < K; j += 32) B[j] = 0.f;
        __syncwarp();
        float norm = 0.f;
        const float cx = out_pos[3 * v], cy = out_pos[3 * v + 1], cz = out_pos[3 * v + 2];
        const float scale = 2.0f / extents[v * extents_stride];
        for (int64_t n0 = b; n0 < e; n0 += 32) {
            const int cnt = (int)min((int64_t)32, e - n0);
            // ---- phase A: lane = neighbour.  Everything that depends only on the neighbour is
            // computed here once: the base cell of its 2x2x2 taps (+ flags for taps clamped onto
            // their twin), the three interpolation fractions and the weighted features.
            float ax = 0.f, ay = 0.f, az = 0.f, nimp_l = 0.f;
            int cellinfo = 0;  // base cell | (x0 == S-1) << 8 | (y0 == S-1) << 9 | (z0 == S-1) << 10
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < cnt) {
                const int p = nidx[n0 + lane];
                nimp_l = nimp ? nimp[n0 + lane] : 1.0f;
                const float imp = inp_importance ? nimp_l * inp_importance[p] : nimp_l;
                f = __ldg(inp_feat + p);
                float x = (inp_pos[3 * (size_t)p] - cx) * scale;
                float y = (inp_pos[3 * (size_t)p + 1] - cy) * scale;
                float z = (inp_pos[3 * (size_t)p + 2] - cz) * scale;
                const float nrm = sqrtf(x * x + y * y + z * z);
                const float amax = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
                const float stretch = amax < 1e-8f ? 0.f : 0.5f * nrm / amax;
                x = fminf(fmaxf((x * stretch + ox + 0.5f) * sm1, 0.f), sm1);
                y = fminf(fmaxf((y * stretch + oy + 0.5f) * sm1, 0.f), sm1);
                z = fminf(fmaxf((z * stretch + oz + 0.5f) * sm1, 0.f), sm1);
                const int x0 = min((int)x, S - 1), y0 = min((int)y, S - 1), z0 = min((int)z, S - 1);
                ax = x - (float)x0;
                ay = y - (float)y0;
                az = z - (float)z0;
                cellinfo = ((z0 * S + y0) * S + x0) | ((x0 == S - 1) << 8) | ((y0 == S - 1) << 9) | ((z0 == S - 1) << 10);
                f.x *= imp;
                f.y *= imp;
                f.z *= imp;
                f.w *= imp;
            }
            // sum of the neighbour importances, in list order like the reference
            for (int j = 0; j < cnt; ++j) norm += __shfl_sync(0xffffffffu, nimp_l, j);
            // the per-neighbour records go through a small per-warp staging area: phase B then needs two
            // broadcast loads per neighbour instead of eight shuffles
            __syncwarp();
            reinterpret_cast<float4*>(stage)[2 * lane] = make_float4(__int_as_float(cellinfo), ax, ay, az);
            reinterpret_cast<float4*>(stage)[2 * lane + 1] = f;
            __syncwarp();
            // ---- phase B: lane = tap x channel
            for (int j = 0; j < cnt; ++j) {
                const float4 rec = reinterpret_cast<const float4*>(stage)[2 * j];
                const float fj = stage[8 * j + 4 + ch];
                const int ci = __float_as_int(rec.x);
                // a "+1" tap that is clamped onto its "+0" twin carries weight 0: skip
                if (!(ci & tapmask)) {
                    const float w = (tx ? rec.y : 1.f - rec.y) * (ty ? rec.z : 1.f - rec.z) * (tz ? rec.w : 1.f - rec.w);
                    B[((ci & 0xff) + tapoff) * 4 + ch] += w * fj;
                }
                __syncwarp();
            }
        }
        // slot bookkeeping with compile-time register indices
#pragma unroll
        for (int i = 0; i < kCcGroup; ++i)
            if (i == g) {
                gv[i] = v;
                gnorm[i] = norm;
            }
        if (++g == kCcGroup) flush();
    }
    if (g > 0) flush();
}

// importance = scale_compat * clamp((1 - d2)^3, 0, 1)
// (CConvAggregationBlock.forward, net_definitions_torch.py:107; window_poly6 common_torch.py:21)
__global__ void __launch_bounds__(256)
importance_kernel(const float* __restrict__ compat, const float* __restrict__ d2, long long n, float* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t = 1.0f - d2[i];
    out[i] = compat[i] * fminf(fmaxf(t * t * t, 0.f), 1.f);
}

void continuous_conv(const float* filters, const float* out_pos, const float* extents, int extents_stride,
                     const float* offset, const float* inp_pos, const float* inp_feat, const float* inp_importance,
                     const int32_t* nidx, const float* nimp, const int64_t* splits, int64_t V, int S, int Cin, int Cout,
                     int normalize, const float* bias, int relu, float* out, cudaStream_t s) {
    if (V == 0) return;
    if (Cin == 4 && S == 4 && ((uintptr_t)inp_feat % 16) == 0) {
        const size_t smem4 = (size_t)(64 * 4 * Cout + kCc4Warps * kCcGroup * 64 * 4 + kCc4Warps * 256) * sizeof(float);
        if (smem4 <= 160 * 1024) {
            ASRB_CUDA(cudaFuncSetAttribute(cconv4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            const unsigned blocks = (unsigned)std::min<size_t>(grid_for(V, kCc4Warps), 148 * 2);
            ProfileScope prof("continuous_conv", s, (double)V * 2.0 * 256 * Cout);
            cconv4_kernel<4><<<blocks, kCc4Warps * 32, smem4, s>>>(filters, out_pos, extents, extents_stride, offset, inp_pos,
                                                                 (const float4*)inp_feat, inp_importance, nidx, nimp,
                                                                 splits, V, Cout, normalize, bias, relu, out);
            ASRB_CHECK_LAUNCH();
            return;
        }
    }
    const size_t smem = (size_t)kCcWarps * S * S * S * Cin * sizeof(float);
    ASRB_REQUIRE(smem <= 200 * 1024, "continuous_conv: kernel_size^3 * in_channels too large for shared memory");
    if (smem > 48 * 1024)
        ASRB_CUDA(cudaFuncSetAttribute(cconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfileScope prof("continuous_conv", s, (double)V * 2.0 * S * S * S * Cin * Cout);
    cconv_kernel<<<grid_for(V, kCcWarps), kCcWarps * 32, smem, s>>>(filters, out_pos, extents, extents_stride, offset,
                                                                   inp_pos, inp_feat, inp_importance, nidx, nimp,
                                                                   splits, V, S, Cin, Cout, normalize, bias, relu, out);
    ASRB_CHECK_LAUNCH();
}

void aggregation_importance(const float* compat, const float* d2, int64_t n, float* out, cudaStream_t s) {
    if (n == 0) return;
    importance_kernel<<<grid_for(n, 256), 256, 0, s>>>(compat, d2, n, out);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
