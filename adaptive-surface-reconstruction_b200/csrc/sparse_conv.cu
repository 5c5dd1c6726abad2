// Generalized sparse convolution over the adaptive grids (K = 55 within a grid,
// K = 9 between grids):
//
//   out[o, :] = sum_{n in row(o)}  imp_n * x[idx_n, :] @ W[slot_n]      (Cin x Cout)
//
// optionally divided by sum_n imp_n (or the row length) and followed by bias +
// ReLU.  Replaces Open3D-ML's `sparse_conv` + `reduce_subarrays_sum` as called
// by SpecialSparseConv.forward (reference models/common_torch.py:95-148).
//
// Open3D fills a dense [K*Cin x 32] matrix per block of 32 outputs and multiplies
// it with the whole filter bank, i.e. it spends K/7.7 ~ 7x the necessary flops on
// these tables.  The B200 design is pair-major ("gather - GEMM - scatter"):
//
//   plan (once per neighbour table, reused by every conv on that grid):
//       stable radix sort of the (row, index, slot) entries by slot, so that each
//       slot owns one contiguous run of pairs and W[slot] is reused by a whole
//       128-pair tile instead of being re-read per output voxel.
//   tile kernel (one CTA = 128 pairs of one slot x TN output channels):
//       cp.async double-buffered staging of the gathered input rows (A, 128 x 32
//       chunk, zero-filled tails) and of the W[slot] chunk (B, 32 x TN) in shared
//       memory; 16 x 16 threads with 8 x (TN/16) register micro-tiles; rows are
//       interleaved (tm + 16 i) and columns split in two 64-wide halves so both
//       shared-memory reads are conflict free; the importance weight is applied
//       to the accumulators per row (it commutes with the contraction) and the
//       tile is added to the output with 16-byte vector reductions
//       (red.global.add.v4.f32).
//   epilogue kernel: normalisation of the importance-normalised channels, bias,
//       ReLU (in place).
//
// The split first convolution of the encoder blocks (conv1a plain | conv1b
// importance-normalised, net_definitions_torch.py:199-210,280-283) runs as ONE
// convolution whose trailing `Cout - imp_col` channels carry the importance.
#include "internal.h"
#include "prims.cuh"
#include "profile.cuh"
#include "sparse_conv.h"

#include <mutex>
#include <vector>

namespace asrb {

constexpr int TM = 128;        // pairs per tile
constexpr int KC = 32;         // input-channel chunk
constexpr int LDA = KC + 4;    // padded row stride of the A tile (floats)
constexpr int kThreads = 256;  // 16 x 16

// ------------------------------------------------------------------ plan
// Pairs are ordered by (row block, kernel slot): rows are cut into blocks of
// 2^kRowBlockShift consecutive output voxels (spatially compact, the grids are
// Morton-ordered within a level) and, inside a block, sorted stably by slot.  A
// block's gathers and reductions then stay L2-resident while its <=55 slot runs
// are processed, and W[slot] is still reused by whole 128-pair tiles.
static int g_row_block_shift = 15;  // dev knob: option conv_row_block_shift
void sparse_conv_row_block_shift(int v) { g_row_block_shift = std::max(10, std::min(v, 24)); }
#define kRowBlockShift g_row_block_shift

// Sort key of an entry: slot-0 entries first (bit `kb`), then (row block, slot).  With the self
// slot first, and every row owning exactly one slot-0 entry (all within-grid tables), the
// slot-0 tiles can STORE their rows — that initialises the whole output, so neither a zero
// fill nor reductions are needed for 1/8 of the pairs (sparse_conv_forward).
__global__ void __launch_bounds__(256)
entry_rows_kernel(const int64_t* __restrict__ splits, long long V, const uint8_t* __restrict__ slot, int kb, int rbs,
                  uint32_t* __restrict__ rows, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                  int* __restrict__ not_one_slot0) {
    const long long v = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    int n0 = 0;
    if (v < V) {
        const int64_t e = splits[v + 1];
        const uint32_t hi = (uint32_t)(v >> rbs) << 8;
        for (int64_t j = splits[v] + sub; j < e; j += 8) {
            const uint32_t k = slot[j];
            rows[j] = (uint32_t)v;
            keys[j] = (k ? (1u << kb) : 0u) | hi | k;
            vals[j] = (uint32_t)j;
            n0 += k == 0;
        }
    }
    n0 += __shfl_xor_sync(0xffffffffu, n0, 1);
    n0 += __shfl_xor_sync(0xffffffffu, n0, 2);
    n0 += __shfl_xor_sync(0xffffffffu, n0, 4);
    if (v < V && sub == 0 && n0 != 1) *not_one_slot0 = 1;
}

// groups in sorted order: g < num_blocks -> (block g, slot 0); then per block the slots 1 .. K-1
__host__ __device__ __forceinline__ void group_block_slot(int g, int K, int num_blocks, int& block, int& slot) {
    if (g < num_blocks) {
        block = g;
        slot = 0;
    } else {
        const int r = g - num_blocks;
        block = r / (K - 1);
        slot = 1 + r % (K - 1);
    }
}

__global__ void __launch_bounds__(256)
gather_pairs_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ rows,
                    const int32_t* __restrict__ idx, long long E, int32_t* __restrict__ p_in,
                    int32_t* __restrict__ p_out) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= E) return;
    const uint32_t e = perm[j];
    p_in[j] = idx[e];
    p_out[j] = (int32_t)rows[e];
}

__device__ __forceinline__ long long lower_bound_u32(const uint32_t* a, long long n, uint32_t k) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (a[mid] < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// first pair of every group: one binary search per thread, whole grid
__global__ void __launch_bounds__(256)
group_begin_kernel(const uint32_t* __restrict__ sorted_key, long long E, int K, int num_blocks, int kb, int G,
                   long long* __restrict__ g_begin) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G) return;
    if (g == G) {
        g_begin[g] = E;
        return;
    }
    int block, slot;
    group_block_slot(g, K, num_blocks, block, slot);
    const uint32_t key = (slot ? (1u << kb) : 0u) | ((uint32_t)block << 8) | (uint32_t)slot;
    g_begin[g] = lower_bound_u32(sorted_key, E, key);
}

// one block: the tile list (slot, first pair, count) for tiles of `tile_rows` pairs.  Prefix over
// the groups: per-thread chunks + a 256-entry scan.
__global__ void __launch_bounds__(256)
tile_list_kernel(int K, int num_blocks, int tile_rows, const long long* __restrict__ g_begin, int* __restrict__ g_tile0,
                 int4* __restrict__ tiles, int* __restrict__ num_tiles) {
    __shared__ int s_part[256];
    const int G = num_blocks * K;
    const int chunk = (G + 255) / 256;
    const int g0 = min(G, (int)threadIdx.x * chunk), g1 = min(G, g0 + chunk);
    int sum = 0;
    for (int g = g0; g < g1; ++g) sum += (int)((g_begin[g + 1] - g_begin[g] + tile_rows - 1) / tile_rows);
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 256; ++i) {
            const int v = s_part[i];
            s_part[i] = t;
            t += v;
        }
        *num_tiles = t;
    }
    __syncthreads();
    int t = s_part[threadIdx.x];
    if (threadIdx.x == 0 && g_tile0) g_tile0[G] = *num_tiles;
    for (int g = g0; g < g1; ++g) {
        const long long b = g_begin[g], e = g_begin[g + 1];
        if (g_tile0) g_tile0[g] = t;
        int block, slot;
        group_block_slot(g, K, num_blocks, block, slot);
        for (long long start = b; start < e; start += tile_rows)
            tiles[t++] = make_int4(slot, (int)start, (int)min((long long)tile_rows, e - start), 0);
    }
}

// Pinned host ints for the plans' "exactly one slot-0 entry per row" flags: slots are handed out from
// pooled pinned pages under a mutex and returned by ~ConvPlan, so a pending plan's slot is never reused.
namespace {
std::mutex g_flag_mutex;
std::vector<int*> g_flag_free;
}  // namespace
int* flag_slot_acquire() {
    std::lock_guard<std::mutex> lock(g_flag_mutex);
    if (g_flag_free.empty()) {
        int* page = nullptr;
        ASRB_CUDA(cudaMallocHost((void**)&page, 256 * sizeof(int)));  // never freed: lives as long as the library
        for (int i = 0; i < 256; ++i) g_flag_free.push_back(page + i);
    }
    int* p = g_flag_free.back();
    g_flag_free.pop_back();
    return p;
}
void flag_slot_release(int* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_flag_mutex);
    g_flag_free.push_back(p);
}

void conv_plan_build(ConvPlan& P, const int32_t* d_idx, const uint8_t* d_slot, const int64_t* d_splits, int64_t V_out,
                     int64_t E, int K, cudaStream_t s) {
    ASRB_REQUIRE(K >= 1 && K <= 256, "sparse_conv: kernel_size must be in [1, 256]");
    ASRB_REQUIRE(E < (int64_t(1) << 31), "sparse_conv: too many neighbour entries");
    ASRB_REQUIRE(V_out < (int64_t(1) << 31), "sparse_conv: too many output rows");
    P.V_out = V_out;
    P.E = E;
    P.K = K;
    const int num_blocks = (int)((V_out + (int64_t(1) << kRowBlockShift) - 1) >> kRowBlockShift);
    const int G = std::max(num_blocks, 1) * K;
    P.max_tiles = (int)((E + TM - 1) / TM) + G;
    P.max_tiles2 = (int)((E + 2 * TM - 1) / (2 * TM)) + G;
    P.tiles2.alloc((size_t)P.max_tiles2, s);
    P.num_tiles2.alloc(1, s);
    P.p_in.alloc((size_t)E, s);
    P.p_out.alloc((size_t)E, s);
    P.perm.alloc((size_t)E, s);
    P.tiles.alloc((size_t)P.max_tiles, s);
    P.num_tiles.alloc(1, s);
    ProfileScope prof("conv_plan_build", s);
    DevBuf<uint32_t> rows((size_t)E, s);
    DevBuf<uint32_t> keys((size_t)E, s);
    P.G = G;
    P.g_begin.alloc((size_t)G + 1, s);
    P.g_tile0.alloc((size_t)G + 1, s);
    DevBuf<long long>& g_begin = P.g_begin;
    DevBuf<int>& g_tile0 = P.g_tile0;
    int bits = 8;
    while (bits < 31 && (int64_t(1) << (bits - 8)) < std::max(num_blocks, 1)) ++bits;
    P.num_blocks = std::max(num_blocks, 1);
    DevBuf<int> flag(1, s);
    ASRB_CUDA(cudaMemsetAsync(flag.get(), 0, sizeof(int), s));
    if (E) {
        entry_rows_kernel<<<grid_for((size_t)V_out * 8, 256), 256, 0, s>>>(d_splits, V_out, d_slot, bits, kRowBlockShift, rows.get(),
                                                                          keys.get(), P.perm.get(), flag.get());
        ASRB_CHECK_LAUNCH();
        sort_pairs_u32_u32(keys.get(), P.perm.get(), (size_t)E, s, bits + 1);
        gather_pairs_kernel<<<grid_for(E, 256), 256, 0, s>>>(P.perm.get(), rows.get(), d_idx, E, P.p_in.get(),
                                                             P.p_out.get());
        ASRB_CHECK_LAUNCH();
    }
    group_begin_kernel<<<grid_for((size_t)G + 1, 256), 256, 0, s>>>(keys.get(), E, K, P.num_blocks, bits, G,
                                                                   g_begin.get());
    ASRB_CHECK_LAUNCH();
    tile_list_kernel<<<1, 256, 0, s>>>(K, P.num_blocks, TM, g_begin.get(), g_tile0.get(), (int4*)P.tiles.get(),
                                       P.num_tiles.get());
    ASRB_CHECK_LAUNCH();
    if (sparse_conv_tc_row_groups() == 2) {
        tile_list_kernel<<<1, 256, 0, s>>>(K, P.num_blocks, 2 * TM, g_begin.get(), nullptr, (int4*)P.tiles2.get(),
                                           P.num_tiles2.get());
        ASRB_CHECK_LAUNCH();
        P.has_tiles2 = true;
    }
    // every row has exactly one slot-0 entry: its tiles are the first `tiles0` of the list and cover
    // every output row once (rows per block / 128, rounded up per block).  The flag travels to a
    // pinned host slot asynchronously and is read at the first convolution (no stall here).
    P.tiles0 = 0;
    P.tiles0_if_flag = 0;
    if (E && K > 1) {
        const int64_t rows_per_block = int64_t(1) << kRowBlockShift;
        for (int b = 0; b < P.num_blocks; ++b) {
            const int64_t r = std::min<int64_t>(rows_per_block, V_out - b * rows_per_block);
            P.tiles0_if_flag += (int)((r + TM - 1) / TM);
        }
        if (!P.flag_host) P.flag_host = flag_slot_acquire();  // owned by the plan until its destructor
        *P.flag_host = 1;
        if (!P.flag_event) ASRB_CUDA(cudaEventCreateWithFlags(&P.flag_event, cudaEventDisableTiming));
        ASRB_CUDA(cudaMemcpyAsync(P.flag_host, flag.get(), sizeof(int), cudaMemcpyDeviceToHost, s));
        ASRB_CUDA(cudaEventRecord(P.flag_event, s));
        P.flag_pending = true;
    }
}

// ------------------------------------------------------------------ tile kernel
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16_cg(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

struct TileArgs {
    const float* x;          // [V_in, Cin]
    const float* w;          // [K, Cin, Cout]
    const int32_t* p_in;     // [E] sorted by slot
    const int32_t* p_out;    // [E]
    const uint32_t* perm;    // [E] original entry of each sorted pair
    const int4* tiles;
    const int* num_tiles;
    const float* imp_in;     // importance per input row (indexed by p_in) or null
    const float* imp_entry;  // importance per original entry (indexed by perm) or null
    float* out;              // [V_out, Cout], zero-initialised, accumulated atomically
    int Cin, Cout;
    int imp_col;             // channels >= imp_col are weighted by the importance
};

template <int TN>
__global__ void __launch_bounds__(kThreads, (TN >= 64) ? 2 : 3)
sparse_conv_tile_kernel(TileArgs a) {
    constexpr int RN = TN / 16;                 // columns per thread: 8, 4 or 2
    constexpr int NV = (RN >= 4) ? RN / 4 : 1;  // float4 column groups per thread
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                 // [2][TM][LDA]
    float* Bs = smem + 2 * TM * LDA;  // [2][KC][TN]
    __shared__ int s_in[TM];
    __shared__ int s_out[TM];
    __shared__ float s_imp[TM];

    if ((int)blockIdx.x >= *a.num_tiles) return;
    const int4 tile = a.tiles[blockIdx.x];
    const int slot = tile.x, start = tile.y, count = tile.z;
    const int n0 = blockIdx.y * TN;
    const int tid = threadIdx.x;
    const int tn = tid & 15, tm = tid >> 4;

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
    __syncthreads();

    const int Cin = a.Cin, Cout = a.Cout;
    const float* wk = a.w + (size_t)slot * Cin * Cout;
    const int nchunks = (Cin + KC - 1) / KC;

    auto load_chunk = [&](int c, int stage) {
        float* As_s = As + stage * TM * LDA;
        float* Bs_s = Bs + stage * KC * TN;
        const int k0 = c * KC;
        // A: TM rows x 8 float4
        for (int i = tid; i < TM * (KC / 4); i += kThreads) {
            const int r = i >> 3, kq = i & 7;
            const int k = k0 + kq * 4;
            const int pin = s_in[r];
            int bytes = 0;
            const float* src = a.x;
            if (pin >= 0 && k < Cin) {
                bytes = min(16, (Cin - k) * 4);
                src = a.x + (size_t)pin * Cin + k;
            }
            cp_async16(As_s + r * LDA + kq * 4, src, bytes);
        }
        // B: KC rows x TN/4 float4
        for (int i = tid; i < KC * (TN / 4); i += kThreads) {
            const int kr = i / (TN / 4), nq = i % (TN / 4);
            const int k = k0 + kr, n = n0 + nq * 4;
            int bytes = 0;
            const float* src = a.w;
            if (k < Cin && n < Cout) {
                bytes = min(16, (Cout - n) * 4);
                src = wk + (size_t)k * Cout + n;
            }
            cp_async16_cg(Bs_s + kr * TN + nq * 4, src, bytes);
        }
    };

    float acc[8][RN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

    load_chunk(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
            load_chunk(c + 1, (c + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* As_s = As + (c & 1) * TM * LDA;
        const float* Bs_s = Bs + (c & 1) * KC * TN;
#pragma unroll 2
        for (int kk = 0; kk < KC; kk += 4) {
            float4 av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(As_s + (tm + 16 * i) * LDA + kk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float bv[RN];
                if constexpr (RN >= 4) {
#pragma unroll
                    for (int g = 0; g < NV; ++g) {
                        const float4 t = *reinterpret_cast<const float4*>(Bs_s + (kk + j) * TN + g * (TN / 2) + tn * 4);
                        bv[4 * g + 0] = t.x;
                        bv[4 * g + 1] = t.y;
                        bv[4 * g + 2] = t.z;
                        bv[4 * g + 3] = t.w;
                    }
                } else {
                    const float2 t = *reinterpret_cast<const float2*>(Bs_s + (kk + j) * TN + tn * 2);
                    bv[0] = t.x;
                    bv[1] = t.y;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av_j = j == 0 ? av[i].x : j == 1 ? av[i].y : j == 2 ? av[i].z : av[i].w;
#pragma unroll
                    for (int q = 0; q < RN; ++q) acc[i][q] = fmaf(av_j, bv[q], acc[i][q]);
                }
            }
        }
        __syncthreads();
    }

    // scatter-add the tile: rows tm + 16 i, column groups of 4 (or 2)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = tm + 16 * i;
        const int o = s_out[r];
        if (o < 0) continue;
        const float imp = s_imp[r];
        float* orow = a.out + (size_t)o * Cout;
        if constexpr (RN >= 4) {
#pragma unroll
            for (int g = 0; g < NV; ++g) {
                const int n = n0 + g * (TN / 2) + tn * 4;
                if (n >= Cout) continue;
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = (n + q >= a.imp_col) ? acc[i][4 * g + q] * imp : acc[i][4 * g + q];
                red_add_v4(orow + n, v[0], v[1], v[2], v[3]);
            }
        } else {
            const int n = n0 + tn * 2;
            if (n >= Cout) continue;
            float v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) v[q] = (n + q >= a.imp_col) ? acc[i][q] * imp : acc[i][q];
            red_add_v2(orow + n, v[0], v[1]);
        }
    }
}

// ------------------------------------------------------------------ importance / epilogue
// out_imp[o] = sum_{e in row(o)} imp[idx[e]]   (gather + reduce_subarrays_sum, common_torch.py:124-128)
__global__ void __launch_bounds__(256)
row_importance_kernel(const float* __restrict__ imp, const int32_t* __restrict__ idx,
                      const int64_t* __restrict__ splits, long long V, float* __restrict__ out) {
    const long long v = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    float s = 0.f;
    if (v < V) {
        const int64_t e = splits[v + 1];
        for (int64_t j = splits[v] + sub; j < e; j += 8) s += idx ? imp[idx[j]] : imp[j];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (v < V && sub == 0) out[v] = s;
}

// in place: channels >= norm_col are divided by the row normaliser (if != 0),
// then bias and ReLU.  normaliser = norm[row] if given else the row length.
__global__ void __launch_bounds__(256)
conv_epilogue_kernel(float* __restrict__ out, long long V, int Cout, int normalize, int norm_col,
                     const float* __restrict__ norm, const int64_t* __restrict__ splits,
                     const float* __restrict__ bias, int relu) {
    const int c4 = Cout >> 2;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V * c4) return;
    const long long row = i / c4;
    const int col = (int)(i - row * c4) * 4;
    float4 v = reinterpret_cast<float4*>(out)[i];
    float e[4] = {v.x, v.y, v.z, v.w};
    if (normalize) {
        const float nrm = norm ? norm[row] : (float)(splits[row + 1] - splits[row]);
        if (nrm != 0.f) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (col + q >= norm_col) e[q] /= nrm;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (bias) e[q] += bias[col + q];
        if (relu) e[q] = fmaxf(e[q], 0.f);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(e[0], e[1], e[2], e[3]);
}

void row_importance(const float* imp, const int32_t* idx, const int64_t* splits, int64_t V, float* out,
                    cudaStream_t s) {
    if (V == 0) return;
    row_importance_kernel<<<grid_for((size_t)V * 8, 256), 256, 0, s>>>(imp, idx, splits, V, out);
    ASRB_CHECK_LAUNCH();
}

template <int TN>
static void launch_tiles(const ConvPlan& P, const TileArgs& a, cudaStream_t s) {
    const size_t smem = (size_t)(2 * TM * LDA + 2 * KC * TN) * sizeof(float);
    static bool configured[64] = {};  // per device (function attributes are per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        ASRB_CUDA(cudaFuncSetAttribute(sparse_conv_tile_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        configured[dev & 63] = true;
    }
    dim3 grid((unsigned)P.max_tiles, (unsigned)((a.Cout + TN - 1) / TN));
    char label[96];
    snprintf(label, sizeof(label), "sparse_conv_tile/fp32 K%d %dx%d E%lld", P.K, a.Cin, a.Cout, (long long)P.E);
    ProfileScope prof(label, s, 2.0 * (double)P.E * a.Cin * a.Cout);
    sparse_conv_tile_kernel<TN><<<grid, kThreads, smem, s>>>(a);
    ASRB_CHECK_LAUNCH();
}

void sparse_conv_forward(const ConvPlan& P, const float* x, const float* w, const float* wp, int Cin, int Cout,
                         const float* imp_in, const float* imp_entry, int imp_col, int normalize, int norm_col, const float* norm,
                         const int64_t* splits, const float* bias, int relu, float* out, cudaStream_t s) {
    ASRB_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, "sparse_conv: channel counts must be multiples of 4");
    if (P.V_out == 0) return;
    if (P.flag_pending) {
        ASRB_CUDA(cudaEventSynchronize(P.flag_event));
        P.tiles0 = *P.flag_host ? 0 : P.tiles0_if_flag;
        P.flag_pending = false;
    }
    const bool store_first = wp && P.E > 0 && P.tiles0 > 0 && sparse_conv_tc_row_groups() == 1;
    if (!store_first) {
        ProfileScope prof("sparse_conv_zero", s);
        ASRB_CUDA(cudaMemsetAsync(out, 0, (size_t)P.V_out * Cout * sizeof(float), s));
    }
    if (P.E > 0 && wp) {
        sparse_conv_tc_tiles(P, x, wp, Cin, Cout, imp_in, imp_entry, imp_col, out, s, store_first);
    } else if (P.E > 0) {
        TileArgs a;
        a.x = x;
        a.w = w;
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
        a.imp_col = (imp_in || imp_entry) ? imp_col : Cout;
        if (Cout > 64) launch_tiles<128>(P, a, s);
        else if (Cout > 32) launch_tiles<64>(P, a, s);
        else launch_tiles<32>(P, a, s);
    }
    if (normalize || bias || relu) {
        ProfileScope prof("sparse_conv_epilogue", s);
        conv_epilogue_kernel<<<grid_for((size_t)P.V_out * (Cout / 4), 256), 256, 0, s>>>(
                out, P.V_out, Cout, normalize, norm_col, norm, splits, bias, relu);
        ASRB_CHECK_LAUNCH();
    }
}

}  // namespace asrb
