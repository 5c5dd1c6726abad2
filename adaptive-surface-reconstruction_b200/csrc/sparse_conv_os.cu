// Output-stationary tensor-core sparse convolution for the within-grid (K = 55)
// tables, used for the convolutions without importance weighting (conv2..4 of
// every block and the decoder blocks: 30 of the 41 within-grid convolutions).
//
// The 55 kernel slots split into 7 "common" ones (self + the 6 same-level face
// neighbours: ~75 % of all pairs, ~80 % filled) and 48 "rare" ones (finer /
// coarser neighbours at level transitions, ~2 % filled each).
//
//   rare slots   -> the pair-major tile kernel (sparse_conv_tc.cu) on the plan of
//                   the rare entries only, reduced into the zeroed output;
//   common slots -> THIS kernel: one CTA owns 128 consecutive output rows, walks
//                   the 7 slots x Cin/16 chunks through the same warp-specialised
//                   mbarrier pipeline (gather -> hi/lo split -> tcgen05.mma
//                   kind::tf32) and keeps the sum over all 7 slots in TMEM.  The
//                   epilogue adds the rare partial sums, applies bias + ReLU and
//                   writes each output row ONCE with coalesced 16-byte stores.
//
// Compared with running everything pair-major this removes 3/4 of the global
// reductions (the measured bottleneck of the small-channel levels), the separate
// epilogue pass, and amortises the per-tile fixed cost over 7x more work.
//
// Accuracy: the tensor pipe truncates on accumulation (see sparse_conv_tc.cu), so
// the chain is kept short: two main accumulators used alternately by slot plus
// one for the 3xTF32 correction terms (3 x n_pad TMEM columns), and the path is
// only taken for Cin <= 128, Cout <= 128 (chain <= 4 x Cin / 8 accumulations).
#include "internal.h"
#include "profile.cuh"
#include "sparse_conv.h"
#include "umma.cuh"

namespace asrb {

namespace {
using umma::bulk_copy_g2s;
using umma::kA_LBO;
using umma::kA_SBO;
using umma::kATileBytes;
using umma::kB_LBO;
using umma::kB_SBO;
using umma::make_desc;
using umma::mbar_arrive;
using umma::mbar_expect_tx;
constexpr int TM = 128;
constexpr int KC = umma::kKC;
constexpr int kSlots = 7;
constexpr int kMaxStages = 4;
constexpr int kPrefetch = 4;
constexpr int kProducerThreads = 128;
constexpr int kThreadsOs = kProducerThreads + 32;
}  // namespace

struct OsArgs {
    const float* x;
    const float* wp;       // packed filters [slot][chunk][hi|lo][n_pad x KC]
    const int32_t* cidx;   // [V][8]
    const float* bias;
    float* out;            // [V, Cout]; holds the rare-slot partial sums when has_rare
    long long V;
    int Cin, Cout, n_pad, stages, relu, has_rare;
};

__global__ void __launch_bounds__(kThreadsOs)
sparse_conv_os_kernel(OsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t b_bytes = (uint32_t)a.n_pad * KC * 4;
    const uint32_t stage_bytes = 2 * kATileBytes + 2 * b_bytes;
    __shared__ uint64_t mbar_full[kMaxStages];
    __shared__ uint64_t mbar_empty[kMaxStages];
    __shared__ uint64_t mbar_acc;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) int s_cidx[TM][8];

    const int tid = threadIdx.x, warp = tid >> 5;
    const long long r0 = blockIdx.x * (long long)TM;
    const int S = a.stages;
    const uint32_t need = 3 * a.n_pad;
    const uint32_t ncols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;

    if (warp == 4) {
        umma::tmem_alloc(&tmem_slot, ncols);
        if ((tid & 31) == 0) {
            for (int i = 0; i < S; ++i) {
                umma::mbar_init(&mbar_full[i], kProducerThreads);
                umma::mbar_init(&mbar_empty[i], 1);
            }
            umma::mbar_init(&mbar_acc, 1);
            umma::fence_barrier_init();
        }
    }
    if (tid < TM) {
        int4 c0 = make_int4(-1, -1, -1, -1), c1 = c0;
        if (r0 + tid < a.V) {
            c0 = reinterpret_cast<const int4*>(a.cidx)[2 * (r0 + tid)];
            c1 = reinterpret_cast<const int4*>(a.cidx)[2 * (r0 + tid) + 1];
        }
        reinterpret_cast<int4*>(&s_cidx[tid][0])[0] = c0;
        reinterpret_cast<int4*>(&s_cidx[tid][0])[1] = c1;
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int Cin = a.Cin;
    const int chunks = (Cin + KC - 1) / KC;
    const int steps = kSlots * chunks;  // step t = (slot k = t / chunks, chunk c = t % chunks)
    const size_t slot_floats = (size_t)chunks * 2 * a.n_pad * KC;

    if (warp < 4) {
        // ------------------------------------------------------------ producers
        const int kq = tid & 3;
        const int rsub = tid >> 2;
        float4 v[kPrefetch][4];
        auto load_step = [&](int t, float4 (&dst)[4]) {
            const int k = t / chunks, c = t - k * chunks;
            const int col = c * KC + kq * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < steps && col < Cin) {
                    const int src = s_cidx[rsub + 32 * j][k];
                    if (src >= 0) dst[j] = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)src * Cin + col));
                }
            }
        };
#pragma unroll
        for (int d = 0; d < kPrefetch; ++d) load_step(d, v[d]);
        for (int t0 = 0; t0 < steps; t0 += kPrefetch) {
#pragma unroll
            for (int d = 0; d < kPrefetch; ++d) {
                const int t = t0 + d;
                if (t < steps) {
                    const int st = t % S, use = t / S;
                    uint8_t* sA_hi = smem + st * stage_bytes;
                    uint8_t* sA_lo = sA_hi + kATileBytes;
                    if (use > 0) umma::mbar_wait(&mbar_empty[st], (use - 1) & 1);
                    if (tid == 0) {
                        const int k = t / chunks, c = t - k * chunks;
                        mbar_expect_tx(&mbar_full[st], 2 * b_bytes);
                        bulk_copy_g2s(sA_lo + kATileBytes, a.wp + k * slot_floats + (size_t)c * 2 * a.n_pad * KC,
                                      2 * b_bytes, &mbar_full[st]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int row = rsub + 32 * j;
                        const float4 x = v[d][j];
                        float4 hi, lo;
                        hi.x = umma::tf32_hi(x.x); hi.y = umma::tf32_hi(x.y);
                        hi.z = umma::tf32_hi(x.z); hi.w = umma::tf32_hi(x.w);
                        lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
                        const uint32_t off = (uint32_t)(row >> 3) * kA_SBO + (uint32_t)kq * kA_LBO + (uint32_t)(row & 7) * 16;
                        *reinterpret_cast<float4*>(sA_hi + off) = hi;
                        *reinterpret_cast<float4*>(sA_lo + off) = lo;
                    }
                    umma::fence_proxy_async();
                    mbar_arrive(&mbar_full[st]);
                    load_step(t + kPrefetch, v[d]);
                }
            }
        }
        // ------------------------------------------------------------ epilogue
        umma::mbar_wait(&mbar_acc, 0);
        umma::tc_fence_after();
        float* T = reinterpret_cast<float*>(smem) + (size_t)warp * 32 * 36;  // per-warp [32 rows][36] transpose buffer
        const int lane = tid & 31;
        for (int n0 = 0; n0 < a.n_pad; n0 += 32) {
            float m0[32], m1[32];
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, m0);
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + a.n_pad + n0, m1);
#pragma unroll
            for (int j = 0; j < 32; ++j) m0[j] += m1[j];
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 2 * a.n_pad + n0, m1);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(T + lane * 36 + j) =
                        make_float4(m0[j] + m1[j], m0[j + 1] + m1[j + 1], m0[j + 2] + m1[j + 2], m0[j + 3] + m1[j + 3]);
            __syncwarp();
            const int cg = lane & 7;
            const int n = n0 + cg * 4;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rl = it * 4 + (lane >> 3);
                const long long row = r0 + warp * 32 + rl;
                if (row < a.V && n < a.Cout) {
                    float4 t = *reinterpret_cast<const float4*>(T + rl * 36 + cg * 4);
                    float4* dst = reinterpret_cast<float4*>(a.out + (size_t)row * a.Cout + n);
                    if (a.has_rare) {
                        const float4 p = *dst;
                        t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
                    }
                    if (a.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n));
                        t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
                    }
                    if (a.relu) {
                        t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
                    }
                    *dst = t;
                }
            }
            __syncwarp();
        }
    } else if ((tid & 31) == 0) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma::make_idesc_tf32(128, a.n_pad);
        const uint32_t t_corr = tmem + 2 * a.n_pad;
        for (int t = 0; t < steps; ++t) {
            const int k = t / chunks, c = t - k * chunks;
            const int st = t % S, use = t / S;
            const uint32_t t_main = tmem + (k & 1) * a.n_pad;  // the two main accumulators alternate by slot
            umma::mbar_wait(&mbar_full[st], use & 1);
            umma::tc_fence_after();
            const uint32_t a_hi = umma::smem_u32(smem + st * stage_bytes), a_lo = a_hi + kATileBytes;
            const uint32_t b_hi = a_lo + kATileBytes, b_lo = b_hi + b_bytes;
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
                const uint32_t oa = ks * 2 * kA_LBO, ob = ks * 2 * kB_LBO;
                const uint64_t dah = make_desc(a_hi + oa, kA_LBO, kA_SBO), dal = make_desc(a_lo + oa, kA_LBO, kA_SBO);
                const uint64_t dbh = make_desc(b_hi + ob, kB_LBO, kB_SBO), dbl = make_desc(b_lo + ob, kB_LBO, kB_SBO);
                umma::mma_tf32(t_main, dah, dbh, idesc, !(k < 2 && c == 0 && ks == 0));
                umma::mma_tf32(t_corr, dal, dbh, idesc, !(t == 0 && ks == 0));
                umma::mma_tf32(t_corr, dah, dbl, idesc, true);
            }
            umma::mma_commit(&mbar_empty[st]);
        }
        umma::mma_commit(&mbar_acc);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tmem, ncols);
}

static bool g_os_enabled = false;
void sparse_conv_os_enable(bool on) { g_os_enabled = on; }
bool sparse_conv_os_enabled() { return g_os_enabled; }

bool sparse_conv_os_supported(const ConvPlan& P, int Cin, int Cout) {
    return g_os_enabled && P.K == 55 && P.rare && P.cidx.size() && Cin <= 128 && Cout <= 128 && Cin % 4 == 0 && Cout % 4 == 0;
}

void sparse_conv_os(const ConvPlan& P, const float* x, const float* wp, int Cin, int Cout, const float* bias, int relu,
                    float* out, cudaStream_t s) {
    const int n_pad = ((Cout + 15) / 16) * 16;
    const bool has_rare = P.rare->E > 0;
    if (has_rare) {
        {
            ProfileScope prof("sparse_conv_zero", s);
            ASRB_CUDA(cudaMemsetAsync(out, 0, (size_t)P.V_out * Cout * sizeof(float), s));
        }
        sparse_conv_tc_tiles(*P.rare, x, wp, Cin, Cout, nullptr, nullptr, Cout, out, s);
    }
    OsArgs a;
    a.x = x;
    a.wp = wp;
    a.cidx = P.cidx.get();
    a.bias = bias;
    a.out = out;
    a.V = P.V_out;
    a.Cin = Cin;
    a.Cout = Cout;
    a.n_pad = n_pad;
    a.relu = relu;
    a.has_rare = has_rare ? 1 : 0;
    const int chunks = (Cin + KC - 1) / KC;
    const size_t stage = 2 * (size_t)kATileBytes + 2 * (size_t)n_pad * KC * 4;
    a.stages = std::max(1, std::min({kMaxStages, kSlots * chunks, (int)((160 * 1024) / stage)}));
    const size_t smem = std::max<size_t>(a.stages * stage, 4 * 32 * 36 * sizeof(float));
    ASRB_CUDA(cudaFuncSetAttribute(sparse_conv_os_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    char label[96];
    snprintf(label, sizeof(label), "sparse_conv_tile/os K55 %dx%d E%lld", Cin, Cout, (long long)P.E_common);
    ProfileScope prof(label, s, 2.0 * (double)P.E_common * Cin * Cout);
    sparse_conv_os_kernel<<<grid_for(P.V_out, TM), kThreadsOs, smem, s>>>(a);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
