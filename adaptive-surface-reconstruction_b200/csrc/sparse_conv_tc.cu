// Tensor-core version of the sparse-convolution tile kernel (see sparse_conv.cu
// for the pair-major plan it runs on).
//
// One CTA = one 128-pair tile of one kernel slot x all output channels
// (N = Cout padded to 16, <= 256):  D[128, N] = A[128, Cin] @ W[slot][Cin, N]
// on tcgen05.mma kind::tf32 with the 3xTF32 split, accumulators in TMEM.
//
// Accuracy note (measured on B200, tools/tc_err.py): the tensor pipe adds each
// MMA into the fp32 accumulator with truncation, a relative bias of about
// -2^-24.3 per accumulation.  The main term A_hi B_hi and the two correction
// terms A_lo B_hi + A_hi B_lo therefore go to SEPARATE accumulators (the
// corrections are 2^-11 of the result, their truncation is invisible) and are
// added in fp32 in the epilogue; the main chain is Cin/8 accumulations long.
//
// Warp-specialised, mbarrier pipeline of up to 4 stages of 16 input channels:
//   * producers (warps 0-3): 4 lanes fetch one 64-byte row chunk (coalesced), keep
//     4 chunks in flight in registers, split into hi/lo and store into the
//     K-major no-swizzle canonical layout (LBO = 144 B so the 16-byte stores of a
//     warp spread evenly over the banks), fence.proxy.async, arrive on `full`.
//   * B (W[slot] chunk, hi and lo): pre-packed once per filter bank in the
//     canonical layout; one `cp.async.bulk` (1-D TMA) per chunk completing its
//     bytes on the same `full` barrier.
//   * MMA issuer (one lane of warp 4): waits `full`, issues the 6 MMAs of the
//     chunk, tcgen05.commit -> `empty` (stage reusable); a last commit -> `acc`.
//   * epilogue (warps 0-3 again): tcgen05.ld (thread = pair = TMEM lane), main +
//     correction, per-row importance on the weighted channels,
//     red.global.add.v4.f32 scatter into the output rows.
#include "internal.h"
#include "profile.cuh"
#include "sparse_conv.h"
#include "umma.cuh"

namespace asrb {

namespace {
constexpr int TM = 128;
constexpr int KC = 16;                           // input channels per pipeline stage (2 MMA k-steps)
constexpr int kMaxStages = 4;
constexpr int kPrefetch = 4;                     // chunks of gathered rows held in registers
constexpr uint32_t kA_LBO = 144;                 // padded: conflict-free staged stores
constexpr uint32_t kA_SBO = (KC / 4) * kA_LBO;   // 576
constexpr uint32_t kATileBytes = 16 * kA_SBO;    // 128 rows -> 9216 B
constexpr uint32_t kB_LBO = 128;
constexpr uint32_t kB_SBO = (KC / 4) * kB_LBO;   // 512
constexpr int kProducerThreads = 128;
constexpr int kThreadsTc = kProducerThreads + 32;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(umma::smem_u32(mbar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         umma::smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(umma::smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
}  // namespace

struct TcArgs {
    const float* x;
    const float* wp;  // packed filters: [slot][chunk][hi|lo][n_pad x KC]
    const int32_t* p_in;
    const int32_t* p_out;
    const uint32_t* perm;
    const int4* tiles;
    const int* num_tiles;
    const float* imp_in;
    const float* imp_entry;
    float* out;
    int Cin, Cout, n_pad, imp_col, stages;
};

// Warp roles: warps 0-3 = producers (gather + hi/lo split) and, after the main
// loop, the epilogue (warp w owns TMEM lanes 32w..32w+31); warp 4 = MMA issuer.
__global__ void __launch_bounds__(kThreadsTc)
sparse_conv_tc_kernel(TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // stage s: A_hi | A_lo | B_hi | B_lo
    const uint32_t b_bytes = (uint32_t)a.n_pad * KC * 4;
    const uint32_t stage_bytes = 2 * kATileBytes + 2 * b_bytes;
    __shared__ uint64_t mbar_full[kMaxStages];   // 128 producer arrivals + B bytes
    __shared__ uint64_t mbar_empty[kMaxStages];  // tcgen05.commit: the MMAs that read the stage are done
    __shared__ uint64_t mbar_acc;                // all MMAs of the tile are done
    __shared__ uint32_t tmem_slot;
    __shared__ int s_in[TM];
    __shared__ int s_out[TM];
    __shared__ float s_imp[TM];

    if ((int)blockIdx.x >= *a.num_tiles) return;
    const int4 tile = a.tiles[blockIdx.x];
    const int slot = tile.x, start = tile.y, count = tile.z;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int S = a.stages;
    // two accumulators: main at column 0, corrections at column n_pad
    const uint32_t ncols = a.n_pad <= 16 ? 32 : a.n_pad <= 32 ? 64 : a.n_pad <= 64 ? 128 : a.n_pad <= 128 ? 256 : 512;

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
        const bool ok = tid < count;
        const int pin = ok ? a.p_in[start + tid] : -1;
        s_in[tid] = pin;
        s_out[tid] = ok ? a.p_out[start + tid] : -1;
        float imp = 1.f;
        if (ok && a.imp_in) imp = a.imp_in[pin];
        if (ok && a.imp_entry) imp *= a.imp_entry[a.perm[start + tid]];
        s_imp[tid] = imp;
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int Cin = a.Cin;
    const int chunks = (Cin + KC - 1) / KC;
    const float* wslot = a.wp + (size_t)slot * chunks * 2 * a.n_pad * KC;

    if (warp < 4) {
        // ------------------------------------------------------------ producers
        const int kq = tid & 3;     // 16-byte column of the chunk
        const int rsub = tid >> 2;  // 0..31: row within a pass (4 passes of 32 rows)
        const float* rowp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pin = s_in[rsub + 32 * j];
            rowp[j] = pin >= 0 ? a.x + (size_t)pin * Cin + kq * 4 : nullptr;
        }
        float4 v[kPrefetch][4];
        auto load_chunk = [&](int c, float4 (&dst)[4]) {
            const int k = c * KC + kq * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rowp[j] && c < chunks && k < Cin) dst[j] = __ldg(reinterpret_cast<const float4*>(rowp[j] + c * KC));
            }
        };
#pragma unroll
        for (int d = 0; d < kPrefetch; ++d) load_chunk(d, v[d]);
        for (int c0 = 0; c0 < chunks; c0 += kPrefetch) {
#pragma unroll
            for (int d = 0; d < kPrefetch; ++d) {
                const int c = c0 + d;
                if (c < chunks) {
                    const int st = c % S, use = c / S;
                    uint8_t* sA_hi = smem + st * stage_bytes;
                    uint8_t* sA_lo = sA_hi + kATileBytes;
                    if (use > 0) umma::mbar_wait(&mbar_empty[st], (use - 1) & 1);
                    if (tid == 0) {
                        mbar_expect_tx(&mbar_full[st], 2 * b_bytes);
                        bulk_copy_g2s(sA_lo + kATileBytes, wslot + (size_t)c * 2 * a.n_pad * KC, 2 * b_bytes,
                                      &mbar_full[st]);
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
                    load_chunk(c + kPrefetch, v[d]);  // refill this register slot
                }
            }
        }
        // ------------------------------------------------------------ epilogue
        umma::mbar_wait(&mbar_acc, 0);
        umma::tc_fence_after();
        // TMEM -> registers (thread = row) -> per-warp transpose through shared memory (the
        // pipeline stages are idle now) so that each RED instruction of a warp covers 4 rows x
        // 128 contiguous bytes (full 32-byte sectors) instead of 32 rows x 16 bytes.
        float* T = reinterpret_cast<float*>(smem) + (size_t)warp * 32 * 36;  // [32 rows][36] per warp
        const int lane = tid & 31;
        for (int n0 = 0; n0 < a.n_pad; n0 += 32) {
            float acc[32], cor[32];
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, acc);
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + a.n_pad + n0, cor);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(T + lane * 36 + j) =
                        make_float4(acc[j] + cor[j], acc[j + 1] + cor[j + 1], acc[j + 2] + cor[j + 2], acc[j + 3] + cor[j + 3]);
            __syncwarp();
            const int cg = lane & 7;
            const int n = n0 + cg * 4;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rl = it * 4 + (lane >> 3);
                const int o = s_out[warp * 32 + rl];
                if (o >= 0 && n < a.Cout) {
                    const float imp = s_imp[warp * 32 + rl];
                    const float4 t = *reinterpret_cast<const float4*>(T + rl * 36 + cg * 4);
                    float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (n + q >= a.imp_col) e[q] *= imp;
                    red_add_v4(a.out + (size_t)o * a.Cout + n, e[0], e[1], e[2], e[3]);
                }
            }
            __syncwarp();
        }
    } else if ((tid & 31) == 0) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma::make_idesc_tf32(128, a.n_pad);
        for (int c = 0; c < chunks; ++c) {
            const int st = c % S, use = c / S;
            umma::mbar_wait(&mbar_full[st], use & 1);
            umma::tc_fence_after();
            const uint32_t a_hi = umma::smem_u32(smem + st * stage_bytes), a_lo = a_hi + kATileBytes;
            const uint32_t b_hi = a_lo + kATileBytes, b_lo = b_hi + b_bytes;
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
                const uint32_t oa = ks * 2 * kA_LBO, ob = ks * 2 * kB_LBO;
                const uint64_t dah = make_desc(a_hi + oa, kA_LBO, kA_SBO), dal = make_desc(a_lo + oa, kA_LBO, kA_SBO);
                const uint64_t dbh = make_desc(b_hi + ob, kB_LBO, kB_SBO), dbl = make_desc(b_lo + ob, kB_LBO, kB_SBO);
                umma::mma_tf32(tmem, dah, dbh, idesc, c > 0 || ks > 0);
                umma::mma_tf32(tmem + a.n_pad, dal, dbh, idesc, c > 0 || ks > 0);
                umma::mma_tf32(tmem + a.n_pad, dah, dbl, idesc, true);
            }
            umma::mma_commit(&mbar_empty[st]);
        }
        umma::mma_commit(&mbar_acc);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tmem, ncols);
}

// [K slots][Cin][Cout] fp32 -> packed hi/lo tiles, see TcArgs::wp
__global__ void __launch_bounds__(256)
pack_conv_filters_kernel(const float* __restrict__ W, int K, int Cin, int Cout, int n_pad, float* __restrict__ out) {
    const int chunks = (Cin + KC - 1) / KC;
    const long long per_slot = (long long)chunks * n_pad * KC;
    const long long total = per_slot * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(i / per_slot);
        const long long r0 = i % per_slot;
        const int c = (int)(r0 / (n_pad * KC));
        const int rem = (int)(r0 % (n_pad * KC));
        const int n = rem / KC, kk = rem % KC;
        const int k = c * KC + kk;
        const float w = (k < Cin && n < Cout) ? W[((size_t)slot * Cin + k) * Cout + n] : 0.f;
        const float hi = umma::tf32_hi(w);
        const size_t tile = ((size_t)slot * chunks + c) * 2 * n_pad * KC;
        // canonical K-major tile: 8-row groups of (KC/4) core matrices of 8 rows x 4 floats
        const size_t off = ((size_t)(n >> 3) * (KC / 4) * 32) + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3);
        out[tile + off] = hi;
        out[tile + (size_t)n_pad * KC + off] = w - hi;
    }
}

size_t packed_conv_filters_floats(int K, int Cin, int Cout) {
    const int n_pad = ((Cout + 15) / 16) * 16;
    return (size_t)K * ((Cin + KC - 1) / KC) * 2 * n_pad * KC;
}

void pack_conv_filters(const float* W, int K, int Cin, int Cout, float* out, cudaStream_t s) {
    const int n_pad = ((Cout + 15) / 16) * 16;
    const long long total = (long long)K * ((Cin + KC - 1) / KC) * n_pad * KC;
    pack_conv_filters_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, s>>>(W, K, Cin, Cout,
                                                                                                        n_pad, out);
    ASRB_CHECK_LAUNCH();
}

void sparse_conv_tc_tiles(const ConvPlan& P, const float* x, const float* wp, int Cin, int Cout, const float* imp_in,
                          const float* imp_entry, int imp_col, float* out, cudaStream_t s) {
    ASRB_REQUIRE(Cout <= 256, "tensor-core sparse conv: out_channels must be <= 256");
    const int n_pad = ((Cout + 15) / 16) * 16;
    TcArgs a;
    a.x = x;
    a.wp = wp;
    a.p_in = P.p_in.get();
    a.p_out = P.p_out.get();
    a.perm = P.perm.get();
    a.tiles = (const int4*)P.tiles.get();
    a.num_tiles = P.num_tiles.get();
    a.imp_in = imp_in;
    a.imp_entry = imp_entry;
    a.out = out;
    a.Cin = Cin;
    a.Cout = Cout;
    a.n_pad = n_pad;
    a.imp_col = (imp_in || imp_entry) ? imp_col : Cout;
    const int chunks = (Cin + KC - 1) / KC;
    const size_t stage = 2 * (size_t)kATileBytes + 2 * (size_t)n_pad * KC * 4;
    a.stages = std::max(1, std::min({n_pad > 128 ? kMaxStages : 3, chunks, (int)((200 * 1024) / stage)}));
    const size_t smem = a.stages * stage;
    ASRB_CUDA(cudaFuncSetAttribute(sparse_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    char label[96];
    snprintf(label, sizeof(label), "sparse_conv_tile/tc K%d %dx%d E%lld", P.K, Cin, Cout, (long long)P.E);
    ProfileScope prof(label, s, 2.0 * (double)P.E * Cin * Cout);
    sparse_conv_tc_kernel<<<(unsigned)P.max_tiles, kThreadsTc, smem, s>>>(a);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
