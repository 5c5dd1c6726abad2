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
//     2-4 chunks in flight in registers, split into hi/lo and store into the
//     K-major no-swizzle canonical layout (8-row groups dense, the four 16-byte
//     column planes 16 * 128 + 32 bytes apart: a quarter-warp's 16-byte stores hit
//     all 32 banks once, umma.cuh), fence.proxy.async, arrive on `full`.
//   * B (W[slot] chunk, hi and lo): pre-packed once per filter bank in the
//     canonical layout; one `cp.async.bulk` (1-D TMA) per chunk completing its
//     bytes on the same `full` barrier.
//   * MMA issuer (one lane of warp 4): waits `full`, issues the 4 MMAs of the
//     chunk — A_hi x [B_hi | B_lo] as ONE MMA of width 2 N (the lo rows follow the
//     hi rows in the stage, the correction accumulator follows the main one in
//     TMEM) and A_lo x B_hi — tcgen05.commit -> `empty`; a last commit -> `acc`.
//   * epilogue (warps 0-3 again): tcgen05.ld (thread = pair = TMEM lane), main +
//     correction, per-row importance on the weighted channels, the pair's output
//     segment staged in shared memory and added to its output row by ONE bulk
//     reduction (cp.reduce.async.bulk.add.f32, TMA engine) — or bulk-STORED for
//     the slot-0 tiles, which come first in the tile list and initialise the output.
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
using umma::red_add_v4;
constexpr int TM = 128;
constexpr int KC = umma::kKC;                    // input channels per pipeline stage (2 MMA k-steps)
constexpr int kMaxStages = 4;
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
    int n_tile;  // output channels handled by one CTA (n_pad, or 128 with blockIdx.y selecting the half)
    int tile0, tile1;  // tiles [tile0, min(tile1, *num_tiles)) of the list
    int store;         // 1: the tiles' rows are written (slot-0 tiles covering every row once), 0: reduced
};

// MT = number of 128-row MMA groups per CTA (tile = 128 * MT pairs).  MT = 2 halves the
// L2 traffic of the filter chunks (one B stage feeds both groups) and the per-tile fixed cost;
// it needs 2 * MT * n_pad TMEM columns, so it is used for n_pad <= 128.
// Warp roles: warps 0 .. 4 MT - 1 = producers (gather + hi/lo split) and, after the main
// loop, the epilogue (warp w: group w / 4, TMEM lanes 32 (w % 4) ..); last warp = MMA issuer.
// PF = chunks of gathered rows a producer keeps in flight in registers (2 for the short
// pipelines of Cin <= 64, which frees registers for 6 resident CTAs per SM; 4 otherwise).
template <int MT, int PF>
__global__ void __launch_bounds__(128 * MT + 32, MT == 2 ? 1 : (PF == 2 ? 5 : 4))
sparse_conv_tc_kernel(TcArgs a) {
    constexpr int kPrefetch = PF;
    constexpr int kProducerThreads = 128 * MT;
    constexpr int kMmaWarp = 4 * MT;
    constexpr int TMC = TM * MT;
    extern __shared__ __align__(1024) uint8_t smem[];
    // stage s: MT x (A_hi | A_lo), then B_hi | B_lo
    const int NT = a.n_tile;
    const int col0 = blockIdx.y * NT;
    const uint32_t b_bytes = (uint32_t)NT * KC * 4;
    const uint32_t stage_bytes = MT * 2 * kATileBytes + 2 * b_bytes;
    __shared__ uint64_t mbar_full[kMaxStages];   // 128 producer arrivals + B bytes
    __shared__ uint64_t mbar_empty[kMaxStages];  // tcgen05.commit: the MMAs that read the stage are done
    __shared__ uint64_t mbar_acc;                // all MMAs of the tile are done
    __shared__ uint32_t tmem_slot;
    __shared__ int s_in[TMC];
    __shared__ int s_out[TMC];
    __shared__ float s_imp[TMC];

    const int tile_id = a.tile0 + (int)blockIdx.x;
    if (tile_id >= min(a.tile1, *a.num_tiles)) return;
    const int4 tile = a.tiles[tile_id];
    const int slot = tile.x, start = tile.y, count = tile.z;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int S = a.stages;
    // per group g two accumulators: main at column 2 g n_pad, corrections at 2 g n_pad + n_pad
    const uint32_t need = 2 * MT * NT;
    const uint32_t ncols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;

    if (warp == kMmaWarp) {
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
    if (tid < TMC) {
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

    if (warp < kMmaWarp) {
        // ------------------------------------------------------------ producers
        const int grp = tid >> 7;           // 128-row group of this thread
        const int kq = tid & 3;             // 16-byte column of the chunk
        const int rsub = (tid & 127) >> 2;  // 0..31: row within a pass (4 passes of 32 rows)
        const float* rowp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pin = s_in[grp * TM + rsub + 32 * j];
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
                    uint8_t* sA_hi = smem + st * stage_bytes + grp * 2 * kATileBytes;
                    uint8_t* sA_lo = sA_hi + kATileBytes;
                    if (use > 0) umma::mbar_wait(&mbar_empty[st], (use - 1) & 1);
                    if (tid == 0) {
                        // hi and lo halves of the chunk, rows [col0, col0 + NT) of each
                        uint8_t* sB = smem + st * stage_bytes + MT * 2 * kATileBytes;
                        const float* src = wslot + (size_t)c * 2 * a.n_pad * KC + (size_t)col0 * KC;
                        mbar_expect_tx(&mbar_full[st], 2 * b_bytes);
                        bulk_copy_g2s(sB, src, b_bytes, &mbar_full[st]);
                        bulk_copy_g2s(sB + b_bytes, src + (size_t)a.n_pad * KC, b_bytes, &mbar_full[st]);
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
        // TMEM -> registers (thread = pair = TMEM lane) -> the pair's output row segment staged
        // in shared memory (the pipeline stages are idle now; row stride NT * 4 + 16 bytes keeps
        // the 16-byte stores conflict free) -> ONE bulk reduction (cp.reduce.async.bulk .add.f32,
        // TMA engine) per pair adds the segment to its output row.  Compared with
        // red.global.add.v4 from registers this takes the scatter off the LSU pipe, which ncu
        // showed to be the limiter of this kernel (l1tex data-pipe wavefronts ~75 %).
        const int lane = tid & 31;
        const int prow = warp * 32 + lane;            // pair index inside the tile
        const uint32_t t_stride = (uint32_t)NT + 4;   // floats
        float* T = reinterpret_cast<float*>(smem) + (size_t)prow * t_stride;
        const uint32_t t_acc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 2 * NT;
        const float imp = s_imp[prow];
        for (int n0 = 0; n0 < NT; n0 += 16) {
            float acc[16], cor[16];
            umma::tmem_ld16(t_acc + n0, acc);
            umma::tmem_ld16(t_acc + NT + n0, cor);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float e[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    e[q] = acc[j + q] + cor[j + q];
                    if (col0 + n0 + j + q >= a.imp_col) e[q] *= imp;
                }
                *reinterpret_cast<float4*>(T + n0 + j) = make_float4(e[0], e[1], e[2], e[3]);
            }
        }
        umma::fence_proxy_async();  // generic-proxy writes of this thread -> visible to the bulk (async proxy) read
        const int o = s_out[prow];
        const int ncol = min(NT, a.Cout - col0);
        if (o >= 0 && ncol > 0) {
            if (a.store) umma::bulk_store(a.out + (size_t)o * a.Cout + col0, T, (uint32_t)ncol * 4);
            else umma::bulk_reduce_add_f32(a.out + (size_t)o * a.Cout + col0, T, (uint32_t)ncol * 4);
        }
        umma::bulk_commit();
        umma::bulk_wait_read();  // the staging rows live in this CTA's shared memory
    } else if ((tid & 31) == 0) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma::make_idesc_tf32(128, NT), idesc2 = umma::make_idesc_tf32(128, 2 * NT);
        const int groups = (MT == 2 && count > TM) ? 2 : 1;  // a half-empty tile skips its second group
        // descriptors = constant part + (address >> 4): a few adds per MMA instead of re-assembly
        const uint64_t da0 = umma::desc_base(kA_LBO, kA_SBO) + (umma::smem_u32(smem) >> 4);
        const uint64_t db0 = umma::desc_base(kB_LBO, kB_SBO) + (umma::smem_u32(smem) >> 4);
        const uint32_t st16 = stage_bytes >> 4, a_tile16 = kATileBytes >> 4;
        int st = 0, ph = 0;
        for (int c = 0; c < chunks; ++c) {
            umma::mbar_wait(&mbar_full[st], ph);
            umma::tc_fence_after();
            const uint64_t a_st = da0 + (uint64_t)(st * st16);
            const uint64_t b_hi0 = db0 + (uint64_t)(st * st16 + MT * 2 * a_tile16);
            for (int g = 0; g < groups; ++g) {
                const uint64_t a_hi0 = a_st + (uint64_t)(g * 2 * a_tile16), a_lo0 = a_hi0 + a_tile16;
                const uint32_t t_main = tmem + g * 2 * NT, t_corr = t_main + NT;
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t oa = (uint64_t)(ks * 2 * (kA_LBO >> 4)), ob = (uint64_t)(ks * 2 * (kB_LBO >> 4));
                    // A_hi x [B_hi | B_lo]: the lo rows follow the hi rows in the stage and the correction
                    // accumulator follows the main one in TMEM, so main and A_hi B_lo are ONE MMA of
                    // width 2 NT (fewer, longer tensor-core instructions; A_hi is read once)
                    umma::mma_tf32(t_main, a_hi0 + oa, b_hi0 + ob, idesc2, c > 0 || ks > 0);
                    umma::mma_tf32_acc(t_corr, a_lo0 + oa, b_hi0 + ob, idesc);
                }
            }
            umma::mma_commit(&mbar_empty[st]);
            if (++st == S) {
                st = 0;
                ph ^= 1;
            }
        }
        umma::mma_commit(&mbar_acc);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) umma::tmem_dealloc(tmem, ncols);
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

// tuning knobs (asr_set_option): pipeline depth for n_pad <= 128 and row groups per CTA
static int g_tc_stages = 2;
static int g_tc_mt = 1;
static int g_tc_ntile = 128;  // output columns per CTA (dev knob: 64 trades a second gather for more CTAs per SM)
void sparse_conv_tc_ntile(int n) { g_tc_ntile = n >= 128 ? 128 : (n >= 64 ? 64 : 32); }
void sparse_conv_tc_tune(int stages, int mt) {
    if (stages > 0) g_tc_stages = std::min(stages, kMaxStages);
    if (mt > 0) g_tc_mt = mt >= 2 ? 2 : 1;
}

int sparse_conv_tc_row_groups() { return g_tc_mt; }

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
                          const float* imp_entry, int imp_col, float* out, cudaStream_t s, bool store_first) {
    ASRB_REQUIRE(Cout <= 256, "tensor-core sparse conv: out_channels must be <= 256");
    const int n_pad = ((Cout + 15) / 16) * 16;
    TcArgs a;
    a.x = x;
    a.wp = wp;
    a.p_in = P.p_in.get();
    a.p_out = P.p_out.get();
    a.perm = P.perm.get();
    a.n_tile = std::min(n_pad, g_tc_ntile);
    while (n_pad % a.n_tile) a.n_tile /= 2;
    const int MT = (g_tc_mt == 2 && P.has_tiles2) ? 2 : 1;
    a.tiles = (const int4*)(MT == 2 ? P.tiles2.get() : P.tiles.get());
    a.num_tiles = MT == 2 ? P.num_tiles2.get() : P.num_tiles.get();
    a.imp_in = imp_in;
    a.imp_entry = imp_entry;
    a.out = out;
    a.Cin = Cin;
    a.Cout = Cout;
    a.n_pad = n_pad;
    a.imp_col = (imp_in || imp_entry) ? imp_col : Cout;
    const int chunks = (Cin + KC - 1) / KC;
    const size_t stage = MT * 2 * (size_t)kATileBytes + 2 * (size_t)a.n_tile * KC * 4;
    a.stages = std::max(1, std::min({g_tc_stages, chunks, (int)((200 * 1024) / stage)}));
    const size_t smem = std::max<size_t>(a.stages * stage, (size_t)128 * MT * (a.n_tile + 4) * sizeof(float));
    char label[96];
    snprintf(label, sizeof(label), "sparse_conv_tile/tc K%d %dx%d E%lld", P.K, Cin, Cout, (long long)P.E);
    ProfileScope prof(label, s, 2.0 * (double)P.E * Cin * Cout);
    const unsigned ny = (unsigned)(n_pad / a.n_tile);
    auto launch = [&](auto kernel, int tiles, int threads) {
        if (tiles <= 0) return;
        ASRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<dim3((unsigned)tiles, ny), threads, smem, s>>>(a);
        ASRB_CHECK_LAUNCH();
    };
    auto run = [&](int t0, int t1, int store) {
        a.tile0 = t0;
        a.tile1 = t1;
        a.store = store;
        if (MT == 2) launch(sparse_conv_tc_kernel<2, 4>, t1 - t0, 288);
        else if (Cin <= 64) launch(sparse_conv_tc_kernel<1, 2>, t1 - t0, 160);
        else launch(sparse_conv_tc_kernel<1, 4>, t1 - t0, 160);
    };
    const int max_tiles = MT == 2 ? P.max_tiles2 : P.max_tiles;
    if (store_first && MT == 1 && P.tiles0 > 0) {
        run(0, P.tiles0, 1);          // slot-0 tiles: write every output row (no zero fill needed)
        run(P.tiles0, max_tiles, 0);  // all other slots: reduce into the written rows
    } else {
        run(0, max_tiles, 0);
    }
}

}  // namespace asrb
