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
//   * A (gathered input rows): 8 lanes fetch one 128-byte row chunk (coalesced),
//     split into hi/lo on the fly and store into the K-major no-swizzle canonical
//     layout with LBO = 144 B so the 16-byte stores of a warp spread evenly over
//     the banks.
//   * B (W[slot] chunk, hi and lo): pre-packed once per filter bank in the
//     canonical layout; one `cp.async.bulk` (1-D TMA) per chunk, completion on an
//     mbarrier with expect_tx.
//   * 2-stage pipeline: the single issuing thread launches the 12 MMAs of a chunk
//     and commits them to the stage's "free" mbarrier; all threads then gather the
//     next chunk into the other stage while the tensor pipe works.
//   * epilogue: tcgen05.ld (thread = pair = TMEM lane), per-row importance on the
//     weighted channels, red.global.add.v4.f32 scatter into the output rows.
#include "internal.h"
#include "profile.cuh"
#include "sparse_conv.h"
#include "umma.cuh"

namespace asrb {

namespace {
constexpr int TM = 128;
constexpr int KC = 32;
constexpr uint32_t kA_LBO = 144;                 // padded: conflict-free staged stores
constexpr uint32_t kA_SBO = 8 * kA_LBO;          // 1152
constexpr uint32_t kATileBytes = 16 * kA_SBO;    // 128 rows -> 18432 B

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(mbar)), "r"(bytes)
                 : "memory");
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
    const float* wp;  // packed filters: [slot][chunk][hi|lo][n_pad x 32]
    const int32_t* p_in;
    const int32_t* p_out;
    const uint32_t* perm;
    const int4* tiles;
    const int* num_tiles;
    const float* imp_in;
    const float* imp_entry;
    float* out;
    int Cin, Cout, n_pad, imp_col;
};

__global__ void __launch_bounds__(128)
sparse_conv_tc_kernel(TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // stage s: A_hi | A_lo | B_hi | B_lo
    const uint32_t b_bytes = (uint32_t)a.n_pad * 128;
    const uint32_t stage_bytes = 2 * kATileBytes + 2 * b_bytes;
    __shared__ uint64_t mbar_full[2];  // B chunk landed (tx bytes)
    __shared__ uint64_t mbar_free[2];  // MMAs that read the stage have completed
    __shared__ uint32_t tmem_slot;
    __shared__ int s_in[TM];
    __shared__ int s_out[TM];
    __shared__ float s_imp[TM];

    if ((int)blockIdx.x >= *a.num_tiles) return;
    const int4 tile = a.tiles[blockIdx.x];
    const int slot = tile.x, start = tile.y, count = tile.z;
    const int tid = threadIdx.x, warp = tid >> 5;
    // two accumulators: main at column 0, corrections at column n_pad
    const uint32_t ncols = a.n_pad <= 16 ? 32 : a.n_pad <= 32 ? 64 : a.n_pad <= 64 ? 128 : a.n_pad <= 128 ? 256 : 512;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        umma::mbar_init(&mbar_full[0], 1);
        umma::mbar_init(&mbar_full[1], 1);
        umma::mbar_init(&mbar_free[0], 1);
        umma::mbar_init(&mbar_free[1], 1);
        umma::fence_barrier_init();
    }
    {
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
    const uint32_t idesc = umma::make_idesc_tf32(128, a.n_pad);
    const int Cin = a.Cin;
    const int chunks = (Cin + KC - 1) / KC;
    const float* wslot = a.wp + (size_t)slot * chunks * 2 * a.n_pad * 32;

    const int kq = tid & 7;    // 16-byte column of the chunk
    const int rsub = tid >> 3;  // 0..15: row within a pass

    for (int c = 0; c < chunks; ++c) {
        const int st = c & 1;
        uint8_t* sA_hi = smem + st * stage_bytes;
        uint8_t* sA_lo = sA_hi + kATileBytes;
        uint8_t* sB = sA_lo + kATileBytes;
        if (c >= 2) umma::mbar_wait(&mbar_free[st], ((c >> 1) - 1) & 1);  // stage no longer read by the tensor pipe
        if (tid == 0) {
            mbar_expect_tx(&mbar_full[st], 2 * b_bytes);
            bulk_copy_g2s(sB, wslot + (size_t)c * 2 * a.n_pad * 32, 2 * b_bytes, &mbar_full[st]);
        }
        // gather: 8 passes of 16 rows; 8 lanes read one 128-byte row chunk
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int row = rsub + 16 * j;
            const int pin = s_in[row];
            const int k = c * KC + kq * 4;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pin >= 0 && k < Cin) v[j] = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)pin * Cin + k));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int row = rsub + 16 * j;
            float4 hi, lo;
            hi.x = umma::tf32_hi(v[j].x); hi.y = umma::tf32_hi(v[j].y);
            hi.z = umma::tf32_hi(v[j].z); hi.w = umma::tf32_hi(v[j].w);
            lo.x = v[j].x - hi.x; lo.y = v[j].y - hi.y; lo.z = v[j].z - hi.z; lo.w = v[j].w - hi.w;
            const uint32_t off = (uint32_t)(row >> 3) * kA_SBO + (uint32_t)kq * kA_LBO + (uint32_t)(row & 7) * 16;
            *reinterpret_cast<float4*>(sA_hi + off) = hi;
            *reinterpret_cast<float4*>(sA_lo + off) = lo;
        }
        umma::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            umma::mbar_wait(&mbar_full[st], (c >> 1) & 1);
            umma::tc_fence_after();
            const uint32_t a_hi = umma::smem_u32(sA_hi), a_lo = umma::smem_u32(sA_lo);
            const uint32_t b_hi = umma::smem_u32(sB), b_lo = b_hi + b_bytes;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t oa = ks * 2 * kA_LBO, ob = ks * 2 * umma::kLBO;
                const uint64_t dah = make_desc(a_hi + oa, kA_LBO, kA_SBO), dal = make_desc(a_lo + oa, kA_LBO, kA_SBO);
                const uint64_t dbh = make_desc(b_hi + ob, umma::kLBO, umma::kSBO),
                               dbl = make_desc(b_lo + ob, umma::kLBO, umma::kSBO);
                umma::mma_tf32(tmem, dah, dbh, idesc, c > 0 || ks > 0);
                umma::mma_tf32(tmem + a.n_pad, dal, dbh, idesc, c > 0 || ks > 0);
                umma::mma_tf32(tmem + a.n_pad, dah, dbl, idesc, true);
            }
            umma::mma_commit(&mbar_free[st]);
        }
    }
    // all MMAs retire in order: waiting for the last commit covers every chunk
    {
        const int last = chunks - 1;
        umma::mbar_wait(&mbar_free[last & 1], (last >> 1) & 1);
    }
    umma::tc_fence_after();

    const int o = s_out[tid];
    const float imp = s_imp[tid];
    float* orow = a.out + (size_t)(o < 0 ? 0 : o) * a.Cout;
    for (int n0 = 0; n0 < a.n_pad; n0 += 32) {
        float v[32], w[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + a.n_pad + n0, w);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += w[j];
        if (o >= 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int n = n0 + j;
                if (n < a.Cout) {
                    float e[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) e[q] = (n + q >= a.imp_col) ? v[j + q] * imp : v[j + q];
                    red_add_v4(orow + n, e[0], e[1], e[2], e[3]);
                }
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

// [K slots][Cin][Cout] fp32 -> packed hi/lo tiles, see TcArgs::wp
__global__ void __launch_bounds__(256)
pack_conv_filters_kernel(const float* __restrict__ W, int K, int Cin, int Cout, int n_pad, float* __restrict__ out) {
    const int chunks = (Cin + 31) / 32;
    const long long per_slot = (long long)chunks * n_pad * 32;
    const long long total = per_slot * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(i / per_slot);
        const long long r0 = i % per_slot;
        const int c = (int)(r0 / (n_pad * 32));
        const int rem = (int)(r0 % (n_pad * 32));
        const int n = rem / 32, kk = rem % 32;
        const int k = c * 32 + kk;
        const float w = (k < Cin && n < Cout) ? W[((size_t)slot * Cin + k) * Cout + n] : 0.f;
        const float hi = umma::tf32_hi(w);
        const size_t tile = ((size_t)slot * chunks + c) * 2 * n_pad * 32;
        const size_t off = ((size_t)(n >> 3) * 256) + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3);
        out[tile + off] = hi;
        out[tile + (size_t)n_pad * 32 + off] = w - hi;
    }
}

size_t packed_conv_filters_floats(int K, int Cin, int Cout) {
    const int n_pad = ((Cout + 15) / 16) * 16;
    return (size_t)K * ((Cin + 31) / 32) * 2 * n_pad * 32;
}

void pack_conv_filters(const float* W, int K, int Cin, int Cout, float* out, cudaStream_t s) {
    const int n_pad = ((Cout + 15) / 16) * 16;
    const long long total = (long long)K * ((Cin + 31) / 32) * n_pad * 32;
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
    const size_t smem = 2 * (2 * (size_t)kATileBytes + 2 * (size_t)n_pad * 128);
    ASRB_CUDA(cudaFuncSetAttribute(sparse_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfileScope prof("sparse_conv_tile", s, 2.0 * (double)P.E * Cin * Cout);
    sparse_conv_tc_kernel<<<(unsigned)P.max_tiles, 128, smem, s>>>(a);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
