// "gx" generalized sparse convolution for the U-Net (reference models/common_torch.py:95-148,
// models/v0/net_definitions_torch.py:123-387 -> Open3D `sparse_conv`):
//
//     out[o, :] = act( sum_{n in row(o)} x[idx_n, :] @ W[slot_n] (/ norm[o]) + bias ) (+ residual[o, :])
//
// B200 design (round 2; replaces the pair-major scatter kernels of round 1 on the model's path):
//
//  * Activations live in HBM in a split-half format between layers: x = hi + lo with two fp16 planes per
//    row (same 4 bytes per element as fp32; |x| < 65504, 22+ significant bits for |x| >= 2^-3, absolute
//    error <= 2^-25 below).  Gathered rows are tensor-core operands AS THEY LIE in memory: 16-byte cp.async
//    copies place them in the K-major 128-byte-swizzled operand tile (no register staging, no conversion).
//    (TMA tile::gather4 does the same with one instruction per 4 rows but is instruction-rate bound on
//    B200 — kept behind the gx_tma_gather option.)
//  * fp32-accurate product on the fp16 pipe (2x the tf32 rate, half the operand bytes of 3xTF32):
//    D += A_hi B_hi + A_lo B_hi + A_hi B_lo; hi*hi products are exact in fp32, the dropped lo*lo term is
//    2^-22 relative.  Filters are packed once per bank as fp16 hi / lo of W * 2^e in the kernel's shared-
//    memory image (one bulk copy per slot and 64-channel chunk).
//  * Output-stationary for the DENSE kernel slots (self + the 6 same-level faces, 73 % of the entries;
//    the 8 child slots of the down tables): one CTA tile = 128 consecutive output rows, accumulators in
//    TMEM across all dense slots, each output row written ONCE with bias / ReLU / normalisation /
//    residual / re-splitting fused — no atomics, no zero fill, no separate epilogue pass, bit-
//    reproducible.  Absent neighbours gather an all-zero row.
//  * The RARE slots (finer / coarser neighbours at level transitions: 27 % of the entries spread over 48
//    slots, present in 87 % of the rows but never dense in any row tile, measured on the bench cloud) run
//    pair-major: entries sorted by (32768-row block, slot) fill 128-pair tiles; their products go to a
//    compact pair buffer in ROW order.  In the output-stationary pass a tile's piece of that buffer is
//    contiguous: one thread streams it through a shared-memory ring with bulk copies and six warps add it
//    up per row, in pair order, into a shared-memory staging tile.
//  * All global traffic of the epilogue is coalesced through that staging tile (rare sums in, the rows'
//    final memory image out, whole 128-byte lines per warp instruction): a thread-per-row epilogue costs
//    32 LSU line transactions per warp instruction and bounded the first version of this kernel
//    (profiles/r2_gx_ablation.txt).
//  * One persistent warp-specialised kernel: 4 gather warps / MMA thread / ring thread / 6 rare-sum
//    warps / 8 epilogue warps (two teams on alternate tiles when N <= 64), 2-8 operand stages, 2-4 TMEM
//    accumulator buffers so the epilogue of a tile overlaps the MMAs of the next ones.
#include <cuda.h>

#include <mutex>
#include <vector>

#include "internal.h"
#include "prims.cuh"
#include "profile.cuh"
#include "spconv_gx.h"
#include "umma.cuh"

namespace asrb {
namespace gx {

static int g_tma_gather = 0;  // dev knob: 1 = gather with TMA tile::gather4 instead of cp.async
void set_tma_gather(int v) { g_tma_gather = v != 0; }
static int g_l1_gather = 0;   // dev knob: gather with cp.async.ca (L1-allocating) — pair with fewer stages
static int g_max_stages = 0;  // dev knob: cap on the pipeline stages (leaves the rest of the 228 KB to L1)
void set_l1_gather(int v) { g_l1_gather = v != 0; }
void set_max_stages(int v) { g_max_stages = v; }
static int g_acc_groups = 0;  // dev knob: main-accumulator groups per TMEM buffer (0 = as many as fit, <= 4)
void set_acc_groups(int g) { g_acc_groups = g; }
// dev knob for ABLATION TIMINGS ONLY (results become garbage): bit 0 no filter copies, 1 no pair-buffer reads,
// 2 no row gathers, 3 no output stores, 4 no MMAs, 5 no generic->async proxy fence before the MMAs
static int g_ablate = 0;
void set_ablate(int v) { g_ablate = v; }
static int g_trace_on = 0;     // dev knob: per-role wait cycles into g_trace (see the kernel)
void set_trace(int v) { g_trace_on = v; }  // 1: the stationary launch, 2: the pair-major launch
static int g_one_team = 0;     // dev knob: one epilogue team even where two fit
void set_one_team(int v) { g_one_team = v != 0; }
static int g_single_tmem = 0;  // dev knob: one TMEM buffer with two accumulator groups for N > 64
void set_single_tmem(int v) { g_single_tmem = v != 0; }

namespace {
constexpr int kRowBlockShift = 15;  // rare entries are grouped by blocks of 2^15 output rows
constexpr int kProducerWarps = 4;
constexpr int kMmaWarp = kProducerWarps;                  // warps 0-3 gather, warp 4 MMA, warps 5-12 epilogue
constexpr int kEpilogueWarps = 8;                         // two per TMEM lane quarter, half of the columns each
constexpr int kRareWarp0 = kProducerWarps + 1 + kEpilogueWarps;  // warps 13-18 sum the rare entries of the tile's rows
constexpr int kRareWarps = 6;
constexpr int kRingWarp = kRareWarp0 + kRareWarps;  // warp 19: one thread feeds the pair-buffer ring
constexpr int kThreads = (kRingWarp + 1) * 32;
constexpr uint32_t kATile = kTM * 128;  // 128 rows x 128 bytes

__device__ int g_overflow_flag = 0;

// dev option gx_trace: cycles that one thread of every role spends in its waits, per CTA (kTraceSlots counters):
//  0 kernel  1 gather warp 0: empty  2 MMA: full  3 MMA: tmem empty  4 epilogue warp 0: rare sums ready  5 tmem full
//  6 its column loop  7 its copy-out  8 its whole tile loop  9 rare warp 0: staging free  10 ring chunk ready
//  11 its whole tile loop  12 ring thread: slot free  13 epilogue warp 4: tmem full  14 MMA whole loop  15 gather warp 0 whole loop
constexpr int kTraceSlots = 16;
__device__ unsigned g_trace[256 * kTraceSlots];
#ifdef GX_TRACE  // `make TRACE=1`: the instrumented build (costs ~20 % even when the option is off)
#define GX_CLOCK() ((unsigned)clock64())
#define GX_TRACING (a.trace != 0)
#define GX_TIMED(cond, slot, stmt)                     \
    do {                                               \
        if (a.trace && (cond)) {                       \
            const unsigned _t0 = (unsigned)clock64();  \
            stmt;                                      \
            tr[slot] += (unsigned)clock64() - _t0;     \
        } else {                                       \
            stmt;                                      \
        }                                              \
    } while (0)
#else
#define GX_CLOCK() 0u
#define GX_TRACING false
#define GX_TIMED(cond, slot, stmt) \
    do {                           \
        stmt;                      \
    } while (0)
#endif

// ------------------------------------------------------------------------------------------ plan
// rare entries (slot >= D) per row
// `row_map` (may be null): the plan covers the table rows row_map[0 .. V) only (this rank's rows of a sharded
// level); local row v = table row row_map[v]
__global__ void __launch_bounds__(256)
rare_count_kernel(const int64_t* __restrict__ splits, const uint8_t* __restrict__ slot, long long V, int D,
                  const int32_t* __restrict__ row_map, int32_t* __restrict__ cnt) {
    const long long v = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    int n = 0;
    if (v < V) {
        const long long gv = row_map ? row_map[v] : v;
        const int64_t e = splits[gv + 1];
        for (int64_t j = splits[gv] + sub; j < e; j += 8) n += slot[j] >= D;
    }
    n += __shfl_xor_sync(0xffffffffu, n, 1);
    n += __shfl_xor_sync(0xffffffffu, n, 2);
    n += __shfl_xor_sync(0xffffffffu, n, 4);
    if (v < V && sub == 0) cnt[v] = n;
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t* __restrict__ p, long long n, int32_t v) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// dense entries -> gather table; rare entries -> (sort key, source row) at their row-major rare position
__global__ void __launch_bounds__(256)
plan_fill_kernel(const int64_t* __restrict__ splits, const int32_t* __restrict__ idx, const uint8_t* __restrict__ slot,
                 long long V, int D, const int32_t* __restrict__ row_map, const int64_t* __restrict__ rare_rs,
                 int32_t* __restrict__ gidx, uint32_t* __restrict__ rare_key, int32_t* __restrict__ rare_src,
                 int32_t* __restrict__ rare_row) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= V) return;
    const long long gv = row_map ? row_map[v] : v;
    int64_t p = rare_rs[v];
    const int64_t e = splits[gv + 1];
    for (int64_t j = splits[gv]; j < e; ++j) {
        const int k = slot[j];
        if (k < D) {
            gidx[((v >> 7) * D + k) * kTM + (v & 127)] = idx[j];
        } else {
            rare_key[p] = ((uint32_t)(v >> kRowBlockShift) << 8) | (uint32_t)k;
            rare_src[p] = idx[j];
            if (rare_row) rare_row[p] = (int32_t)gv;  // pair-final form: the destination is the (table) row itself
            ++p;
        }
    }
}

__global__ void __launch_bounds__(256) iota_u32_kernel(uint32_t* __restrict__ p, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

__device__ __forceinline__ long long lower_bound_u32(const uint32_t* a, long long n, uint32_t k) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (a[mid] < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// first sorted rare entry of every (row block, slot) group; g = block * K + slot
__global__ void __launch_bounds__(256)
group_begin_kernel(const uint32_t* __restrict__ sorted_key, long long R, int K, int G, long long* __restrict__ g_begin) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G) return;
    if (g == G) {
        g_begin[g] = R;
        return;
    }
    const uint32_t key = ((uint32_t)(g / K) << 8) | (uint32_t)(g % K);
    g_begin[g] = lower_bound_u32(sorted_key, R, key);
}

// one block: 128-pair tiles of the groups (slot, first entry, count)
__global__ void __launch_bounds__(256)
pair_tile_list_kernel(int K, int G, const long long* __restrict__ g_begin, int4* __restrict__ tiles,
                      int* __restrict__ num_tiles) {
    __shared__ int s_part[256];
    const int chunk = (G + 255) / 256;
    const int g0 = min(G, (int)threadIdx.x * chunk), g1 = min(G, g0 + chunk);
    int sum = 0;
    for (int g = g0; g < g1; ++g) sum += (int)((g_begin[g + 1] - g_begin[g] + kTM - 1) / kTM);
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
    for (int g = g0; g < g1; ++g) {
        const long long b = g_begin[g], e = g_begin[g + 1];
        for (long long start = b; start < e; start += kTM)
            tiles[t++] = make_int4(g % K, (int)start, (int)min((long long)kTM, e - start), 0);
    }
}

// per pair tile: gather rows and destinations of its 128 lanes
__global__ void __launch_bounds__(kTM)
pair_tile_fill_kernel(const int4* __restrict__ tiles, const int* __restrict__ num_tiles, const uint32_t* __restrict__ perm,
                      const int32_t* __restrict__ rare_src, const int32_t* __restrict__ out_row, int32_t zero_row,
                      int32_t* __restrict__ pt_slot, int32_t* __restrict__ pt_gidx, int32_t* __restrict__ pt_out) {
    const int t = blockIdx.x;
    if (t >= *num_tiles) return;
    const int4 tile = tiles[t];
    const int r = threadIdx.x;
    int src = zero_row, dst = -1;
    if (r < tile.z) {
        const uint32_t p = perm[tile.y + r];
        src = rare_src[p];
        dst = out_row ? out_row[p] : (int32_t)p;
    }
    pt_gidx[(size_t)t * kTM + r] = src;
    pt_out[(size_t)t * kTM + r] = dst;
    if (r == 0) pt_slot[t] = tile.x;
}

}  // namespace

Plan::~Plan() {
    if (R_event) {
        if (!finished) cudaEventSynchronize(R_event);  // the async copy into the slot must have landed
        cudaEventDestroy(R_event);
    }
    pinned_slot_release(R_host);
}

void plan_begin(Plan& P, const int32_t* d_idx, const uint8_t* d_slot, const int64_t* d_splits, int64_t V, int64_t V_in,
                int64_t E, int K, int mode, const int32_t* d_row_map, cudaStream_t s) {
    P.d_row_map = d_row_map;
    ASRB_REQUIRE(K >= 1 && K <= 255, "gx plan: kernel_size must be in [1, 255]");
    ASRB_REQUIRE(V < (int64_t(1) << 31) - 2 && V_in < (int64_t(1) << 31) - 2 && E < (int64_t(1) << 31),
                 "gx plan: table too large");
    P.V_in = V_in;
    ASRB_REQUIRE(mode == kModeStationary || mode == kModePairFinal, "gx plan: bad mode");
    P.mode = mode;
    P.V = V;
    P.E = E;
    P.K = K;
    // dense slots: the self slot and the six same-level faces of a within-grid table (grid.cpp:102-124), the
    // eight child slots of a down table (inverted up table, grid.cpp:217-242); none for the pair-final form
    P.D = mode == kModePairFinal ? 0 : (K == 55 ? 7 : std::min(K, 8));
    P.T = (V + kTM - 1) / kTM;
    P.d_idx = d_idx;
    P.d_slot = d_slot;
    P.d_splits = d_splits;
    P.rare_rs.alloc((size_t)V + 1, s);
    ProfileScope prof("gx_plan_build", s);
    DevBuf<int32_t> cnt((size_t)std::max<int64_t>(V, 1), s);
    if (V > 0) {
        rare_count_kernel<<<grid_for((size_t)V * 8, 256), 256, 0, s>>>(d_splits, d_slot, V, P.D, d_row_map, cnt.get());
        ASRB_CHECK_LAUNCH();
    }
    exclusive_sum_i32_to_i64(cnt.get(), P.rare_rs.get(), (size_t)V, s);
    if (!P.R_host) P.R_host = pinned_slot_acquire();
    if (!P.R_event) ASRB_CUDA(cudaEventCreateWithFlags(&P.R_event, cudaEventDisableTiming));
    ASRB_CUDA(cudaMemcpyAsync(P.R_host, P.rare_rs.get() + V, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    ASRB_CUDA(cudaEventRecord(P.R_event, s));
    P.finished = false;
}

void plan_finish(Plan& P, cudaStream_t s) {
    if (P.finished) return;
    ASRB_CUDA(cudaEventSynchronize(P.R_event));
    P.R = *P.R_host;
    const int64_t V = P.V, R = P.R;
    const int32_t zero_row = (int32_t)P.V_in;  // the all-zero row that follows the input tensor's rows
    ProfileScope prof("gx_plan_build", s);
    PhaseTimer pt(s);
    if (P.D > 0) {
        const size_t n = (size_t)P.T * P.D * kTM;
        P.gidx.alloc(std::max<size_t>(n, 1), s);
        if (n) {
            fill_i32_kernel<<<grid_for(n, 256), 256, 0, s>>>(P.gidx.get(), (long long)n, zero_row);
            ASRB_CHECK_LAUNCH();
        }
    }
    DevBuf<uint32_t> key((size_t)std::max<int64_t>(R, 1), s), perm((size_t)std::max<int64_t>(R, 1), s);
    P.rare_in.alloc((size_t)std::max<int64_t>(R, 1), s);
    DevBuf<int32_t>& src = P.rare_in;
    DevBuf<int32_t> rows;
    if (P.mode == kModePairFinal) rows.alloc((size_t)std::max<int64_t>(R, 1), s);
    if (V > 0) {
        plan_fill_kernel<<<grid_for((size_t)V, 256), 256, 0, s>>>(P.d_splits, P.d_idx, P.d_slot, V, P.D, P.d_row_map,
                                                                 P.rare_rs.get(), P.gidx.get(), key.get(), src.get(),
                                                                 P.mode == kModePairFinal ? rows.get() : nullptr);
        ASRB_CHECK_LAUNCH();
    }
    pt.lap("gx plan: fill", (long long)R);
    const int num_blocks = (int)std::max<int64_t>((V + (int64_t(1) << kRowBlockShift) - 1) >> kRowBlockShift, 1);
    const int G = num_blocks * P.K;
    P.max_pair_tiles = (int)(R / kTM) + G + 1;
    P.pt_slot.alloc((size_t)P.max_pair_tiles, s);
    P.pt_gidx.alloc((size_t)P.max_pair_tiles * kTM, s);
    P.pt_out.alloc((size_t)P.max_pair_tiles * kTM, s);
    P.pt_count.alloc(1, s);
    if (R > 0) {
        iota_u32_kernel<<<grid_for((size_t)R, 256), 256, 0, s>>>(perm.get(), R);
        ASRB_CHECK_LAUNCH();
        int bits = 8;
        while (bits < 31 && (int64_t(1) << (bits - 8)) < num_blocks) ++bits;
        sort_pairs_u32_u32(key.get(), perm.get(), (size_t)R, s, bits);
        pt.lap("gx plan: sort", (long long)R);
        DevBuf<long long> g_begin((size_t)G + 1, s);
        group_begin_kernel<<<grid_for((size_t)G + 1, 256), 256, 0, s>>>(key.get(), R, P.K, G, g_begin.get());
        ASRB_CHECK_LAUNCH();
        DevBuf<int4> tiles((size_t)P.max_pair_tiles, s);
        pair_tile_list_kernel<<<1, 256, 0, s>>>(P.K, G, g_begin.get(), tiles.get(), P.pt_count.get());
        ASRB_CHECK_LAUNCH();
        pt.lap("gx plan: tile list", (long long)G);
        pair_tile_fill_kernel<<<(unsigned)P.max_pair_tiles, kTM, 0, s>>>(
                tiles.get(), P.pt_count.get(), perm.get(), src.get(), P.mode == kModePairFinal ? rows.get() : nullptr,
                zero_row, P.pt_slot.get(), P.pt_gidx.get(), P.pt_out.get());
        ASRB_CHECK_LAUNCH();
    } else {
        ASRB_CUDA(cudaMemsetAsync(P.pt_count.get(), 0, sizeof(int), s));
    }
    pt.lap("gx plan: tile fill", (long long)P.max_pair_tiles);
    P.d_idx = nullptr;
    P.d_slot = nullptr;
    P.d_splits = nullptr;
    P.finished = true;
}

// ------------------------------------------------------------------------------------------ filters
namespace {
// byte offset of element (row n, half k) of a [rows x 64 halves] K-major tile with the 128-byte swizzle
// (Swizzle<3,4,3>: the 16-byte chunk index is XORed with the row index modulo 8)
__host__ __device__ __forceinline__ uint32_t sw128_offset(int n, int k) {
    return (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ n) & 7) << 4) + (k & 7) * 2);
}

// Operand layout rule.  64-channel stages: an A stage = hi tile + lo tile (2 x 16 KB), B = [B_hi | B_lo] (2 N x 128
// bytes).  32-channel stages ("c32"): ONE tile per operand whose 128-byte rows hold [32 hi halves | 32 lo halves].
// c32 is used for 32 input channels.  (Measured and rejected: c32 stages also for N = 128, which makes room for a
// second staging tile so that the rare sums run ahead of the epilogue — twice the stages, barrier round trips and
// MMA instructions cost more than that overlap gains: level-1 128->128 0.95 -> 1.05 ms, level-2 256->128 0.35 -> 0.48 ms.)
__host__ __device__ __forceinline__ bool use_c32(int Cin, int N) { return (void)N, Cin == 32; }

__global__ void __launch_bounds__(256)
pack_filters_kernel(const float* __restrict__ W, int K, int Cin, int Cout, int col0, int ncols, int N, float scale,
                    uint8_t* __restrict__ out) {
    const bool c32 = use_c32(Cin, N);
    const int chunks = c32 ? Cin / 32 : Cin / 64;
    const int kc = c32 ? 32 : 64;
    const size_t chunk_bytes = (size_t)(c32 ? 1 : 2) * N * 128;
    const long long total = (long long)K * chunks * N * kc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % kc);
        long long r = i / kc;
        const int n = (int)(r % N);
        r /= N;
        const int c = (int)(r % chunks);
        const int slot = (int)(r / chunks);
        const int k = c * kc + kk;
        const float w = n < ncols ? W[((size_t)slot * Cin + k) * Cout + col0 + n] * scale : 0.f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        uint8_t* tile = out + ((size_t)slot * chunks + c) * chunk_bytes;
        if (c32) {  // one tile: halves 0..31 = hi, 32..63 = lo (same row layout as a 32-channel activation row)
            *reinterpret_cast<__half*>(tile + sw128_offset(n, kk)) = hi;
            *reinterpret_cast<__half*>(tile + sw128_offset(n, kk + 32)) = lo;
        } else {  // hi tile, then lo tile
            *reinterpret_cast<__half*>(tile + sw128_offset(n, kk)) = hi;
            *reinterpret_cast<__half*>(tile + (size_t)N * 128 + sw128_offset(n, kk)) = lo;
        }
    }
}
}  // namespace

static int padded_n(int ncols) { return ((ncols + 15) / 16) * 16; }

size_t packed_filter_bytes(int K, int Cin, int ncols) {
    const int N = padded_n(ncols);
    return (size_t)K * (Cin / 32) * N * 128;  // the same for both operand layouts
}

void pack_filters(const float* W, int K, int Cin, int Cout, int col0, int ncols, int scale_exp, void* out,
                  cudaStream_t s) {
    ASRB_REQUIRE(Cin == 32 || (Cin >= 64 && Cin % 64 == 0), "gx: in_channels must be 32 or a multiple of 64");
    ASRB_REQUIRE(ncols >= 1 && ncols <= 256 && col0 >= 0 && col0 + ncols <= Cout, "gx: bad output column range");
    const int N = padded_n(ncols);
    const long long total = (long long)K * Cin * N;
    pack_filters_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, s>>>(
            W, K, Cin, Cout, col0, ncols, N, ldexpf(1.f, scale_exp), (uint8_t*)out);
    ASRB_CHECK_LAUNCH();
}

// ------------------------------------------------------------------------------------------ kernel
namespace {

__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const CUtensorMap* tmap, int col, int r0, int r1, int r2,
                                            int r3, uint32_t mbar) {
    asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
            "%4, %5, %6}], [%7];" ::"r"(smem_dst),
            "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(mbar)
            : "memory");
}

// kind::f16 (fp16 x fp16 -> fp32), A and B K-major, M = 128
__device__ __forceinline__ uint32_t idesc_f16(int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
}
// K-major, 128-byte swizzle: LBO field 1 (unused), SBO = 1024 bytes (8 rows), version 1, layout type 2
__device__ __forceinline__ uint64_t sw128_desc_base() {
    return ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

enum { kKindStationary = 0, kKindPairBuf = 1, kKindPairFinal = 2 };

struct KArgs {
    // work
    const int32_t* gidx;       // [steps][128]
    const int32_t* tile_slot;  // pair kinds: slot of each tile
    const int32_t* out_pos;    // pair kinds: destination of each lane
    const int* num_tiles_dev;  // pair kinds
    int num_tiles;             // stationary kind
    int D;                     // steps per stationary tile
    int V;
    // input: rows of x_pitch halves, hi / lo planes at columns a_hi / a_lo (also the tensor-map coordinates)
    const __half* x;
    int x_pitch;
    int a_hi, a_lo, chunks;
    int tma_gather, l1_gather, ablate, trace;
    // filters
    const uint8_t* wp;
    unsigned long long slot_bytes;
    uint32_t chunk_bytes;
    // MMA shape
    int N, ncat, stages;
    int G, nbuf, by_slot;  // accumulator groups per TMEM buffer, buffers, group = kernel slot (importance variant)
    int teams;             // epilogue teams (1 or 2)
    int nrb, rare_lp;      // staging tiles (1 or 2), lanes per pair row of the rare warps (power of two >= N / 4)
    int ring_slots, ring_pairs;  // pair-buffer ring: chunks of ring_pairs pair rows
    // importance variant (conv1b of SpecialSparseConv): every gathered row is weighted by imp[input row]
    const float* imp;
    const int32_t* rare_in;  // [R] input row of every rare entry (row-major rare order)
    int zero_row;
    uint32_t stage_bytes, tx_bytes;
    // epilogue
    float wscale;
    const float* bias;
    const float* norm;
    int relu, ncols;
    const int64_t* rare_rs;
    const int32_t* row_map;  // stationary kind: table row of every local row, or null
    const float* pairbuf;  // read (stationary)
    float* pair_out;       // written (pair-buffer kind)
    __half* out_h;
    int out_pitch, out_hi, out_lo;
    float* out_f;
    int out_f_pitch;
    const __half* res;
    int res_pitch, res_hi, res_lo;
};

__device__ __forceinline__ void split_store8(__half* hi_p, __half* lo_p, const float* v, int& overflow) {
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float x = v[j];
        if (fabsf(x) > 65504.f) {
            overflow = 1;
            x = copysignf(65504.f, x);
        }
        h[j] = __float2half_rn(x);
        l[j] = __float2half_rn(x - __half2float(h[j]));
    }
    *reinterpret_cast<uint4*>(hi_p) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo_p) = *reinterpret_cast<const uint4*>(l);
}

template <bool C32, int KIND>
__global__ void __launch_bounds__(kThreads, 1) gx_conv_kernel(const __grid_constant__ CUtensorMap tmap, const KArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_full[8], bar_empty[8], bar_tfull[4], bar_tempty[4], bar_rfull[2], bar_rempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar_ring_full[4], bar_ring_empty[4];
    __shared__ long long s_rs[2][kTM + 1];      // rare-segment bounds of the tile's rows
    __shared__ float s_w[kRareWarps][32];       // importance of the pairs of a ring chunk
    const uint32_t sbase = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;  // operand tiles need 1024-byte alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = a.stages;
    const int N = a.N;
    // TMEM.  The tensor pipe TRUNCATES every addition into the fp32 accumulator (a relative bias of about
    // -2^-24.3 per MMA, round 1's tools/tc_err.py, the same on kind::f16).  Therefore (i) the corrections
    // A_lo B_hi + A_hi B_lo — 2^-11 of the result, their own truncation invisible — go to their own accumulator
    // and never lengthen the main chain, and (ii) the main chain is cut into G groups (round robin over the
    // (slot, 64-channel chunk) stages) that the epilogue adds in fp32 with round-to-nearest: G x [main N | corr N]
    // columns per buffer, two buffers when they fit.  The importance variant uses one group per kernel slot and
    // applies the per-row importance to each group in the epilogue (exact fp32 scaling).
    const int G = a.G, nbuf = a.nbuf;
    const int accw = 2 * N * G;
    uint32_t ncols_alloc = 32;
    while ((int)ncols_alloc < nbuf * accw) ncols_alloc <<= 1;
    const int ntiles = KIND == kKindStationary ? a.num_tiles : *a.num_tiles_dev;

    if (warp == kMmaWarp) {
        umma::tmem_alloc(&tmem_slot, ncols_alloc);
        if (lane == 0) {
            for (int i = 0; i < S; ++i) {
                // TMA gather: one arrive.expect_tx; cp.async gather: one arrival per producer thread + the filters' one
                umma::mbar_init(&bar_full[i], a.tma_gather ? 1 : kProducerWarps * 32 + 1);
                umma::mbar_init(&bar_empty[i], 1);
            }
            for (int i = 0; i < 4; ++i) {
                umma::mbar_init(&bar_tfull[i], 1);
                umma::mbar_init(&bar_tempty[i], kEpilogueWarps / a.teams);
            }
            for (int i = 0; i < 2; ++i) {
                // shared-memory hand-overs between warps: EVERY lane arrives after its own accesses (an elected lane behind
                // __syncwarp is ordered too, but compute-sanitizer's racecheck does not follow that chain)
                umma::mbar_init(&bar_rfull[i], kRareWarps * 32);
                umma::mbar_init(&bar_rempty[i], kEpilogueWarps / a.teams * 32);
            }
            for (int i = 0; i < 4; ++i) {
                umma::mbar_init(&bar_ring_full[i], 1);
                umma::mbar_init(&bar_ring_empty[i], kRareWarps * 32);
            }
            umma::fence_barrier_init();
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int steps_per_tile = KIND == kKindStationary ? a.D : 1;
    unsigned tr[kTraceSlots];
#pragma unroll
    for (int i = 0; i < kTraceSlots; ++i) tr[i] = 0;
    const unsigned t_kernel0 = GX_CLOCK();

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------------ gather producers
        // Default: 16-byte cp.async (LDGSTS) straight into the swizzled K-major operand tiles.  Lane l of a warp
        // instruction copies chunk l % 8 of row 4 i + l / 8: four whole 128-byte lines per instruction, written to
        // ((chunk ^ row) & 7) * 16 inside the row's 128-byte slot (Swizzle<3,4,3>, what the MMA descriptor expects).
        // Absent neighbours (the zero row) are zero-filled without a memory read (src-size 0).
        // The TMA form (tile::gather4, one instruction per 4 rows x 128 bytes; dev option gx_tma_gather) is kept
        // for reference: measured on B200 it sustains only ~1 instruction per ~100 cycles and SM (1.3-1.6 TB/s
        // chip-wide whatever the shape), 4-5x below what the L2 delivers to the LSU path.
        int st = 0, ph = 0;
        const int rl = lane >> 3, cj = lane & 7;
        for (int tile = blockIdx.x; a.tma_gather && warp == 0 && tile < ntiles; tile += gridDim.x) {
            for (int sidx = 0; sidx < steps_per_tile; ++sidx) {
                const long long step = KIND == kKindStationary ? (long long)tile * a.D + sidx : tile;
                const int slot = KIND == kKindStationary ? sidx : a.tile_slot[tile];
                const uint8_t* wsrc = a.wp + (size_t)slot * a.slot_bytes;
                {
                    const int4 r = *reinterpret_cast<const int4*>(a.gidx + step * kTM + lane * 4);
                    for (int c = 0; c < a.chunks; ++c) {
                        umma::mbar_wait(&bar_empty[st], ph ^ 1);
                        const uint32_t stage = sbase + (uint32_t)st * a.stage_bytes;
                        const uint32_t full = umma::smem_u32(&bar_full[st]);
                        if (lane == 0) umma::mbar_arrive_expect_tx(&bar_full[st], a.tx_bytes);
                        __syncwarp();
                        tma_gather4(stage + lane * 512, &tmap, a.a_hi + c * 64, r.x, r.y, r.z, r.w, full);
                        if (!C32) tma_gather4(stage + kATile + lane * 512, &tmap, a.a_lo + c * 64, r.x, r.y, r.z, r.w, full);
                        if (lane == 0) {
                            const uint32_t bdst = stage + (C32 ? 1u : 2u) * kATile;
                            asm volatile(
                                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bdst),
                                    "l"(wsrc + (size_t)c * a.chunk_bytes), "r"(a.chunk_bytes), "r"(full)
                                    : "memory");
                        }
                        if (++st == S) {
                            st = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
        if (!a.tma_gather) {
            // (tile, step) pairs of this CTA as one sequence, the gather rows of step q + 1 loaded while step q is
            // being issued (the index loads are dependent global loads: without the prefetch they cost a full
            // memory latency per step on the critical path of all four producer warps)
            const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            const long long nq = (long long)my_tiles * steps_per_tile;
            auto step_of = [&](long long q, int& slot) -> long long {
                const int tile = (int)blockIdx.x + (int)(q / steps_per_tile) * (int)gridDim.x;
                const int sidx = (int)(q % steps_per_tile);
                if (KIND == kKindStationary) {
                    slot = sidx;
                    return (long long)tile * a.D + sidx;
                }
                slot = a.tile_slot[tile];
                return tile;
            };
            int g_next[8], slot_next = 0;
            if (nq > 0) {
                const long long step = step_of(0, slot_next);
#pragma unroll
                for (int i = 0; i < 8; ++i) g_next[i] = a.gidx[step * kTM + warp * 32 + 4 * i + rl];
            }
            for (long long q = 0; q < nq; ++q) {
                int g[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = g_next[i];
                const int slot = slot_next;
                if (q + 1 < nq) {
                    const long long step = step_of(q + 1, slot_next);
#pragma unroll
                    for (int i = 0; i < 8; ++i) g_next[i] = a.gidx[step * kTM + warp * 32 + 4 * i + rl];
                }
                const uint8_t* wsrc = a.wp + (size_t)slot * a.slot_bytes;
                // the 8 rows this lane copies from: rows warp * 32 + 4 i + lane / 8
                const __half* src[8];
                uint32_t dst[8], nbytes[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = warp * 32 + 4 * i + rl;
                    src[i] = a.x + (size_t)g[i] * a.x_pitch;
                    nbytes[i] = g[i] == a.zero_row ? 0u : 16u;
                    dst[i] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((cj ^ r) & 7) << 4));
                }
                // column of this lane's 16-byte piece: c32 rows are [32 hi | 32 lo] of the chunk's 32 channels
                const int col_hi = C32 ? (cj < 4 ? a.a_hi + cj * 8 : a.a_lo + (cj - 4) * 8) : a.a_hi + cj * 8;
                const int col_lo = a.a_lo + cj * 8, cstep = C32 ? 32 : 64;
                for (int c = 0; c < a.chunks; ++c) {
                    GX_TIMED(warp == 0 && lane == 0, 1, umma::mbar_wait(&bar_empty[st], ph ^ 1));
                    const uint32_t stage = sbase + (uint32_t)st * a.stage_bytes;
                    if (warp == 0 && lane == 0 && (a.ablate & 1)) {
                        umma::mbar_arrive(&bar_full[st]);
                    } else if (warp == 0 && lane == 0) {
                        const uint32_t bdst = stage + (C32 ? 1u : 2u) * kATile;
                        umma::mbar_arrive_expect_tx(&bar_full[st], a.chunk_bytes);
                        asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bdst),
                                "l"(wsrc + (size_t)c * a.chunk_bytes), "r"(a.chunk_bytes), "r"(umma::smem_u32(&bar_full[st]))
                                : "memory");
                    }
                    if (a.ablate & 4) {
                    } else if (a.l1_gather) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            umma::cp_async16_ca(stage + dst[i], src[i] + col_hi + c * cstep, nbytes[i]);
                        if (!C32) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                umma::cp_async16_ca(stage + kATile + dst[i], src[i] + col_lo + c * cstep, nbytes[i]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            umma::cp_async16_cg(stage + dst[i], src[i] + col_hi + c * cstep, nbytes[i]);
                        if (!C32) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                umma::cp_async16_cg(stage + kATile + dst[i], src[i] + col_lo + c * cstep, nbytes[i]);
                        }
                    }
                    umma::cp_async_arrive_noinc(&bar_full[st]);
                    if (++st == S) {
                        st = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idN = idesc_f16(N), id2N = idesc_f16(2 * N);
            const uint64_t dbase = sw128_desc_base();
            int st = 0, ph = 0, it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int buf = it % nbuf;
                GX_TIMED(true, 3, umma::mbar_wait(&bar_tempty[buf], ((it / nbuf) & 1) ^ 1));
                umma::tc_fence_after();
                const uint32_t acc0 = tmem + (uint32_t)(buf * accw);
                uint32_t started = 0;  // bit g: group g's accumulators hold this tile's data (else the MMA overwrites)
                int sc = 0;
                for (int sidx = 0; sidx < steps_per_tile; ++sidx) {
                    for (int c = 0; c < a.chunks; ++c, ++sc) {
                        GX_TIMED(true, 2, umma::mbar_wait(&bar_full[st], ph));
                        // the cp.async writes (generic proxy, acquired through the barrier) -> visible to the MMA's
                        // operand reads (async proxy); this thread has no loads of its own in flight, so it is cheap
                        if (!(a.ablate & 32)) umma::fence_proxy_async();
                        umma::tc_fence_after();
                        const uint32_t stage = sbase + (uint32_t)st * a.stage_bytes;
                        const int g = a.by_slot ? sidx : (sc % G);
                        const uint32_t acc = acc0 + (uint32_t)(g * 2 * N), cor = acc + (uint32_t)N;
                        uint32_t first = (started >> g) & 1u, firstc = first;
                        started |= 1u << g;
                        if (a.ablate & 16) {
                        } else if (C32) {
                            // one A tile [hi 32 | lo 32], one B tile [hi 32 | lo 32]: k-steps 0,1 = hi, 2,3 = lo
                            const uint64_t da = dbase + (stage >> 4), db = dbase + ((stage + kATile) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                mma_f16(acc, da + 2 * ks, db + 2 * ks, idN, first);              // hi hi
                                first = 1;
                                mma_f16(cor, da + 2 * (ks + 2), db + 2 * ks, idN, firstc);       // lo hi
                                firstc = 1;
                                mma_f16(cor, da + 2 * ks, db + 2 * (ks + 2), idN, 1);            // hi lo
                            }
                        } else {
                            const uint64_t dah = dbase + (stage >> 4), dal = dbase + ((stage + kATile) >> 4);
                            const uint64_t dbh = dbase + ((stage + 2 * kATile) >> 4);
                            const uint64_t dbl = dbh + (uint64_t)((N * 128) >> 4);
                            if (a.ncat) {
                                // A_hi x [B_hi | B_lo] as ONE MMA of width 2 N (the lo rows follow the hi rows, the
                                // correction accumulator follows the main one), then A_lo x B_hi onto the corrections
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    mma_f16(acc, dah + 2 * ks, dbh + 2 * ks, id2N, first);
                                    first = 1;
                                    mma_f16(cor, dal + 2 * ks, dbh + 2 * ks, idN, 1);
                                }
                            } else {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    mma_f16(acc, dah + 2 * ks, dbh + 2 * ks, idN, first);
                                    first = 1;
                                    mma_f16(cor, dal + 2 * ks, dbh + 2 * ks, idN, firstc);
                                    firstc = 1;
                                    mma_f16(cor, dah + 2 * ks, dbl + 2 * ks, idN, 1);
                                }
                            }
                        }
                        umma::mma_commit(&bar_empty[st]);
                        if (++st == S) {
                            st = 0;
                            ph ^= 1;
                        }
                    }
                }
                umma::mma_commit(&bar_tfull[buf]);
            }
        }
        __syncwarp();
    } else if (warp < kRareWarp0) {
        // ------------------------------------------------------------------ epilogue
        // 8 warps: warp e reads TMEM lane quarter e % 4 (thread = output row / pair) and handles half e / 4 of the
        // columns (<= 64).  All global traffic of the epilogue is COALESCED through a shared-memory staging tile
        // (the LSU processes one 128-byte line per cycle: a thread-per-row store or load costs 32 line transactions
        // per warp instruction, and those — not the gathers — bounded the first version of this kernel, see
        // profiles/r2_gx_ablation.txt):
        //   1. the rare warps have summed the row's pair-buffer segment into the staging row -> registers;
        //   2. accumulator groups are added, the epilogue applied, the result written back to the staging row in its
        //      final memory image ([hi halves | lo halves], or fp32);
        //   3. the two warps of the lane quarter copy their 32 rows out, consecutive lanes = consecutive 16 bytes.
        const int e = warp - (kMmaWarp + 1);
        const int q = warp & 3;  // TMEM lane quarter this warp may access (warps 5..12 -> 1,2,3,0,1,2,3,0)
        const int L = q * 32 + lane;
        // one team: the two warps of a quarter split the columns of a tile; two teams (N <= 64, two staging tiles):
        // warps e < 4 take the even tiles of this CTA, the others the odd ones, a thread = a whole row — two tiles
        // in the epilogue at a time, which is what a latency-bound epilogue (pair kinds: one MMA step per tile) needs
        const int teams = a.teams;
        const int team = teams == 2 ? e >> 2 : 0;
        const int half = teams == 2 ? 0 : e >> 2;
        const int tp = half * 32 + lane;  // thread of the quarter's warp pair
        const int cph = ((N / 16 + 1) / 2) * 16;  // columns of half 0 (multiple of 16)
        const int c0 = teams == 2 ? 0 : (half ? cph : 0), c1 = teams == 2 ? N : (half ? N : cph);
        const bool has_rare = KIND == kKindStationary && a.rare_rs != nullptr && !(a.ablate & 2);
        const uint32_t ldr = 4u * (uint32_t)N + 16u;  // staging row pitch in bytes (+16: conflict-free thread-per-row access)
        uint8_t* const stage_gen = smem_raw + (sbase - umma::smem_u32(smem_raw)) + (size_t)S * a.stage_bytes;
        const int nrb = a.nrb;
        int overflow = 0;
        int it = 0;
        // output row (final kinds) / pair-buffer row of this thread's lane in a tile, and its normaliser
        auto load_row = [&](int tile, long long& row, float& nrm) {
            row = -1;
            if (KIND == kKindStationary) {
                const long long v = (long long)tile * kTM + L;
                if (v < a.V) row = a.row_map ? a.row_map[v] : v;
            } else {
                row = a.out_pos[(size_t)tile * kTM + L];
            }
            nrm = (KIND != kKindPairBuf && a.norm && row >= 0) ? a.norm[row] : 0.f;
        };
        // copy-out: every warp copies what it wrote itself — its 32 rows x its columns [c0, c0 + nv) — so the staging
        // tile needs no cross-warp synchronisation.  16-byte chunk ch of the warp's part of a row -> offset in the
        // staging row, destination base and pitch:
        const int nv = KIND == kKindPairBuf ? c1 - c0 : max(0, min(c1, a.ncols) - c0);  // valid columns of this half
        const int CW = nv >> 2;                                                       // chunks per row
        auto copy_geometry = [&](int ch, size_t& pitch, uint32_t& soff) -> uint8_t* {
            if (KIND == kKindPairBuf) {
                pitch = (size_t)N * 4;
                soff = 4 * c0 + ch * 16;
                return reinterpret_cast<uint8_t*>(a.pair_out + c0) + ch * 16;
            }
            if (a.out_f) {
                pitch = (size_t)a.out_f_pitch * 4;
                soff = 4 * c0 + ch * 16;
                return reinterpret_cast<uint8_t*>(a.out_f + c0) + ch * 16;
            }
            // split-half rows: columns c0 + 8 cc .. + 7 of plane `plane` = block cc / 2 of this half, second part of it if cc is odd
            const int hc = CW >> 1, plane = ch >= hc ? 1 : 0, cc = ch - plane * hc;
            pitch = (size_t)a.out_pitch * 2;
            soff = 4 * c0 + (cc >> 1) * 64 + plane * 32 + (cc & 1) * 16;
            return reinterpret_cast<uint8_t*>(a.out_h + (plane ? a.out_lo : a.out_hi) + c0) + cc * 16;
        };
        // chunks per row a power of two (the rule): this lane copies chunk lane % CW of every (32 / CW)-th row
        int cw_shift = -1;
        for (int sh = 0; sh <= 5; ++sh)
            if ((1 << sh) == CW) cw_shift = sh;
        const int co_r0 = cw_shift >= 0 ? lane >> cw_shift : 0, co_rstep = cw_shift >= 0 ? 32 >> cw_shift : 1;
        size_t co_pitch = 0;
        uint32_t co_src = 0;
        uint8_t* const co_base = copy_geometry(lane & (CW - 1), co_pitch, co_src);
        long long row_n = -1;
        float nrm_n = 0.f;
        const int tstep = teams * (int)gridDim.x;
        it = team;
        if ((int)blockIdx.x + team * (int)gridDim.x < ntiles) load_row(blockIdx.x + team * gridDim.x, row_n, nrm_n);
        for (int tile = blockIdx.x + team * gridDim.x; tile < ntiles; tile += tstep, it += teams) {
            const int buf = it % nbuf, rb = it % nrb;
            uint8_t* const Rb = stage_gen + (size_t)rb * kTM * ldr;
            uint8_t* const myrow = Rb + (size_t)L * ldr;
            // destination row and normaliser of this tile (loaded one tile ahead), the next tile's on their way
            const long long row = row_n;
            const float nrm = nrm_n;
            if (tile + tstep < ntiles) load_row(tile + tstep, row_n, nrm_n);
            // groups of this tile that hold data; importance variant: the weight of every dense slot's row
            const int used = a.by_slot ? steps_per_tile : min(G, steps_per_tile * a.chunks);
            float wgt[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) wgt[g] = 1.f;
            if (KIND == kKindStationary && a.imp) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (g < used) {
                        const int src = a.gidx[((size_t)tile * a.D + g) * kTM + L];
                        wgt[g] = src == a.zero_row ? 0.f : a.imp[src];
                    }
                }
            }
            // the sums of the rows' rare entries are in the staging tile (columns [c0, c1) of a row are touched by
            // this thread only: they are read block by block below and replaced by the block's results)
            if (has_rare) GX_TIMED(e == 0 && lane == 0, 4, umma::mbar_wait(&bar_rfull[rb], (it / nrb) & 1));
            GX_TIMED((e == 0 || e == 4) && lane == 0, e == 0 ? 5 : 13, umma::mbar_wait(&bar_tfull[buf], (it / nbuf) & 1));
            const unsigned t_cb0 = GX_CLOCK();
            umma::tc_fence_after();
            const uint32_t t_acc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * accw);
            for (int cb = 0; cb < 8; ++cb) {
                const int n0 = c0 + 16 * cb;
                if (n0 >= c1) break;  // warp-uniform
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
                __syncwarp();  // the TMEM loads are warp-collective: reconverge after the row-dependent code below
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (g < used) {
                        uint32_t mc[32];  // main and correction columns: both loads in flight, one wait
                        umma::tmem_ld16_issue(t_acc + g * 2 * N + n0, *reinterpret_cast<uint32_t(*)[16]>(&mc[0]));
                        umma::tmem_ld16_issue(t_acc + g * 2 * N + N + n0, *reinterpret_cast<uint32_t(*)[16]>(&mc[16]));
                        umma::tmem_ld_wait32(mc);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            v[j] = fmaf(__uint_as_float(mc[j]) + __uint_as_float(mc[16 + j]), wgt[g], v[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], a.wscale, 0.f);
                if (n0 >= a.ncols && KIND != kKindPairBuf) continue;  // warp-uniform
                float4* const blk = reinterpret_cast<float4*>(myrow + 4 * n0);  // this block's 64 bytes of the staging row
                if (KIND == kKindPairBuf) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) blk[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    continue;
                }
                if (has_rare) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 t = blk[j];
                        v[4 * j] += t.x;
                        v[4 * j + 1] += t.y;
                        v[4 * j + 2] += t.z;
                        v[4 * j + 3] += t.w;
                    }
                }
                const int nvalid = min(16, a.ncols - n0);  // multiple of 8
                if (a.norm && nrm != 0.f) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] /= nrm;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (a.bias && j < nvalid) v[j] += a.bias[n0 + j];
                    if (a.relu) v[j] = fmaxf(v[j], 0.f);
                }
                if (a.res && row >= 0) {
                    const __half* rh = a.res + (size_t)row * a.res_pitch + a.res_hi + n0;
                    const __half* rl = a.res + (size_t)row * a.res_pitch + a.res_lo + n0;
                    for (int j0 = 0; j0 < nvalid; j0 += 8) {
                        const uint4 uh = *reinterpret_cast<const uint4*>(rh + j0);
                        const uint4 ul = *reinterpret_cast<const uint4*>(rl + j0);
                        const __half* hh = reinterpret_cast<const __half*>(&uh);
                        const __half* ll = reinterpret_cast<const __half*>(&ul);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j0 + j] += __half2float(hh[j]) + __half2float(ll[j]);
                    }
                }
                if (a.out_f) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) blk[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                    // split-half image of the block, in place: [16 hi halves | 16 lo halves]
                    __half* oh = reinterpret_cast<__half*>(blk);
                    int of = 0;
                    split_store8(oh, oh + 16, v, of);
                    split_store8(oh + 8, oh + 24, v + 8, of);
                    if (row >= 0) overflow |= of;
                }
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_tempty[buf]);
            __syncwarp();
            const unsigned t_co0 = GX_CLOCK();
            if (GX_TRACING && e == 0 && lane == 0) tr[6] += t_co0 - t_cb0;
            // copy-out: consecutive lanes = consecutive 16-byte chunks of a row
            const int myrow_dst = (a.ablate & 8) ? -1 : (int)row;
            if (cw_shift >= 0) {
#pragma unroll 4
                for (int r = co_r0; r < 32; r += co_rstep) {
                    const int drow = __shfl_sync(0xffffffffu, myrow_dst, r);
                    if (drow >= 0)
                        *reinterpret_cast<uint4*>(co_base + (size_t)drow * co_pitch) =
                                *reinterpret_cast<const uint4*>(Rb + (size_t)(q * 32 + r) * ldr + co_src);
                }
            } else if (CW > 0) {  // chunks per row not a power of two
                for (int i0 = 0; i0 < 32 * CW; i0 += 32) {
                    const int idx = i0 + lane;
                    const int r = min(idx / CW, 31);
                    const int drow = __shfl_sync(0xffffffffu, myrow_dst, r);
                    if (idx < 32 * CW && drow >= 0) {
                        size_t pitch;
                        uint32_t soff;
                        uint8_t* const base = copy_geometry(idx - r * CW, pitch, soff);
                        *reinterpret_cast<uint4*>(base + (size_t)drow * pitch) =
                                *reinterpret_cast<const uint4*>(Rb + (size_t)(q * 32 + r) * ldr + soff);
                    }
                }
            }
            if (GX_TRACING && e == 0 && lane == 0) tr[7] += GX_CLOCK() - t_co0;
            if (has_rare) umma::mbar_arrive(&bar_rempty[rb]);
        }
        if (GX_TRACING && e == 0 && lane == 0) tr[8] = GX_CLOCK() - t_kernel0;
        if (overflow) atomicOr(&g_overflow_flag, 1);
    } else if (warp == kRingWarp) {
        // ------------------------------------------------------------------ pair-buffer ring producer (one thread of
        // its own warp: as a second lane of the MMA warp it was starved by that warp's barrier polling)
        if (lane == 0 && KIND == kKindStationary && a.rare_rs != nullptr && !(a.ablate & 2)) {
            // pair-buffer ring producer: the tiles' rare segments, chunk by chunk, with bulk copies
            const uint32_t ldr = 4u * (uint32_t)N + 16u;
            const uint32_t ring = sbase + (uint32_t)S * a.stage_bytes + (uint32_t)a.nrb * kTM * ldr;
            const int CP = a.ring_pairs, RS = a.ring_slots;
            const uint32_t slot_bytes = (uint32_t)CP * N * 4;
            int rsl = 0, rph = 0;
            long long n0 = 0, n1 = 0;
            if ((int)blockIdx.x < ntiles) {
                const long long t0 = (long long)blockIdx.x * kTM;
                n0 = a.rare_rs[t0];
                n1 = a.rare_rs[min(t0 + kTM, (long long)a.V)];
            }
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const long long tp0 = n0, tp1 = n1;
                if (tile + (int)gridDim.x < ntiles) {  // bounds of the next tile while this one streams
                    const long long t0 = (long long)(tile + gridDim.x) * kTM;
                    n0 = a.rare_rs[t0];
                    n1 = a.rare_rs[min(t0 + kTM, (long long)a.V)];
                }
                for (long long c = tp0; c < tp1; c += CP) {
                    const uint32_t bytes = (uint32_t)min((long long)CP, tp1 - c) * (uint32_t)N * 4u;
                    GX_TIMED(true, 12, umma::mbar_wait(&bar_ring_empty[rsl], rph ^ 1));
                    umma::mbar_arrive_expect_tx(&bar_ring_full[rsl], bytes);
                    asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + (uint32_t)rsl * slot_bytes),
                            "l"(a.pairbuf + (size_t)c * N), "r"(bytes), "r"(umma::smem_u32(&bar_ring_full[rsl]))
                            : "memory");
                    if (++rsl == RS) {
                        rsl = 0;
                        rph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ rare-entry sums (stationary kind)
        // The products of a tile's rare entries are ONE contiguous piece of the pair buffer (row order).  The ring
        // warp streams it through a small ring of shared-memory chunks with bulk copies (cp.async.bulk: no
        // LSU transactions, deep memory-level parallelism from one instruction); a group of LP = N / 4 lanes of
        // these warps owns every (128 / LP)-th row and adds their pairs per row IN PAIR ORDER (the order of the
        // single-threaded sum it replaces: results are bit-identical) from the ring into the staging tile.
        const bool has_rare = KIND == kKindStationary && a.rare_rs != nullptr && !(a.ablate & 2);
        if (has_rare) {
            const int rw = warp - kRareWarp0;
            const int rt = rw * 32 + lane;  // thread of the rare-sum warps
            const int LP = a.rare_lp;
            const int NG = kRareWarps * 32 / LP;  // lane groups; group g owns rows g, g + NG, g + 2 NG, ... so that the
            const int g = rt / LP;              // pairs of one ring chunk (consecutive rows) spread over all groups
            const int lig = lane & (LP - 1);
            const bool colok = 4 * lig < N;
            const uint32_t ldr = 4u * (uint32_t)N + 16u;
            uint8_t* const stage_gen = smem_raw + (sbase - umma::smem_u32(smem_raw)) + (size_t)S * a.stage_bytes;
            const int nrb = a.nrb;
            const float* const ring = reinterpret_cast<const float*>(stage_gen + (size_t)nrb * kTM * ldr);
            const int CP = a.ring_pairs, RS = a.ring_slots;
            int rsl = 0, rph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int rb = it % nrb;
                uint8_t* const Rb = stage_gen + (size_t)rb * kTM * ldr;
                const long long t0 = (long long)tile * kTM;
                long long* const rs = s_rs[it & 1];  // rare-segment bounds of the tile's rows (double-buffered: one barrier per tile)
                if (rt <= kTM) rs[rt] = a.rare_rs[min(t0 + rt, (long long)a.V)];
                asm volatile("bar.sync 5, %0;" ::"n"(kRareWarps * 32) : "memory");
                const long long tp0 = rs[0];
                const int total = (int)(rs[kTM] - tp0);  // pairs of the tile; all positions below are relative to tp0
                float wnext = (a.imp && lane < total) ? a.imp[a.rare_in[tp0 + lane]] : 0.f;  // first chunk's importances
                GX_TIMED(rw == 0 && lane == 0, 9, umma::mbar_wait(&bar_rempty[rb], ((it / nrb) & 1) ^ 1));
                int cr = g;  // current row
                int p = (int)(rs[cr] - tp0), pe = (int)(rs[cr + 1] - tp0);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                uint8_t* dst = Rb + (size_t)cr * ldr + 16 * lig;
                const uint32_t dstep = (uint32_t)NG * ldr;
                auto finish_row = [&]() {
                    if (colok) *reinterpret_cast<float4*>(dst) = acc;
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    cr += NG;
                    dst += dstep;
                    if (cr < kTM) {
                        p = (int)(rs[cr] - tp0);
                        pe = (int)(rs[cr + 1] - tp0);
                    }
                };
                for (int cb = 0; cb < total; cb += CP) {
                    const int ce = min(cb + CP, total);
                    if (a.imp) {  // importance of the chunk's pairs (one per lane; CP <= 32), loaded one chunk ahead
                        __syncwarp();
                        s_w[rw][lane] = wnext;
                        __syncwarp();
                        wnext = ce + lane < total ? a.imp[a.rare_in[tp0 + ce + lane]] : 0.f;
                    }
                    GX_TIMED(rw == 0 && lane == 0, 10, umma::mbar_wait(&bar_ring_full[rsl], rph));
                    const float* const chunk = ring + (size_t)rsl * CP * N + (colok ? 4 * lig : 0);
                    while (cr < kTM && p < ce) {
                        const int hi = min(pe, ce);
                        const float* src = chunk + (p - cb) * N;
                        if (a.imp) {
                            for (; p < hi; ++p, src += N) {
                                const float4 t = *reinterpret_cast<const float4*>(src);
                                const float w = s_w[rw][p - cb];
                                acc.x = fmaf(t.x, w, acc.x);
                                acc.y = fmaf(t.y, w, acc.y);
                                acc.z = fmaf(t.z, w, acc.z);
                                acc.w = fmaf(t.w, w, acc.w);
                            }
                        } else {
                            for (; p + 1 < hi; p += 2, src += 2 * N) {  // both loads in flight, added in pair order
                                const float4 t = *reinterpret_cast<const float4*>(src);
                                const float4 u = *reinterpret_cast<const float4*>(src + N);
                                acc.x = (acc.x + t.x) + u.x;
                                acc.y = (acc.y + t.y) + u.y;
                                acc.z = (acc.z + t.z) + u.z;
                                acc.w = (acc.w + t.w) + u.w;
                            }
                            if (p < hi) {
                                const float4 t = *reinterpret_cast<const float4*>(src);
                                acc.x += t.x;
                                acc.y += t.y;
                                acc.z += t.z;
                                acc.w += t.w;
                                ++p;
                            }
                        }
                        if (p < pe) break;  // the row continues in the next chunk
                        finish_row();
                    }
                    umma::mbar_arrive(&bar_ring_empty[rsl]);
                    if (++rsl == RS) {
                        rsl = 0;
                        rph ^= 1;
                    }
                }
                while (cr < kTM) finish_row();  // rows without (further) pairs
                umma::mbar_arrive(&bar_rfull[rb]);
            }
        }
    }
    if (GX_TRACING && lane == 0) {
        // every role's lane 0 holds its own counters; slot ownership is disjoint
        const unsigned now = GX_CLOCK() - t_kernel0;
        unsigned* out = g_trace + (blockIdx.x & 255) * kTraceSlots;
        if (warp == 0) { out[0] = now; out[1] = tr[1]; out[15] = now; }
        if (warp == kMmaWarp) { out[2] = tr[2]; out[3] = tr[3]; out[14] = now; }
        if (warp == kMmaWarp + 1) { out[4] = tr[4]; out[5] = tr[5]; out[6] = tr[6]; out[7] = tr[7]; out[8] = tr[8]; }
        if (warp == kMmaWarp + 5) out[13] = tr[13];
        if (warp == kRareWarp0) { out[9] = tr[9]; out[10] = tr[10]; out[11] = now; }
        if (warp == kRingWarp) out[12] = tr[12];
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) umma::tmem_dealloc(tmem, ncols_alloc);
}

// ---------------------------------------------------------------- tensor map
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    if (!fn) throw Error(kCudaError, "gx: cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

// rows of `pitch` halves; box = 64 halves (128 bytes) x 1 row; the gather instruction moves 4 boxes
CUtensorMap make_row_map(const H2View& x) {
    ASRB_REQUIRE(x.p != nullptr && x.rows >= 1, "gx: null activation view");
    ASRB_REQUIRE(((uintptr_t)x.p & 15) == 0 && (x.pitch * 2) % 16 == 0 && x.pitch >= 64,
                 "gx: activation rows must be 16-byte aligned and at least 64 halves wide");
    CUtensorMap m;
    const cuuint64_t gdim[2] = {(cuuint64_t)x.pitch, (cuuint64_t)x.rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)x.pitch * 2};
    const cuuint32_t box[2] = {64, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)x.p, gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(kCudaError, "gx: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <bool C32, int KIND>
void launch(const CUtensorMap& tmap, const KArgs& k, int grid, size_t smem, cudaStream_t s) {
    ASRB_CUDA(cudaFuncSetAttribute(gx_conv_kernel<C32, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (k.l1_gather)  // no more shared memory than the stages need: the remainder of the 228 KB serves as L1
        ASRB_CUDA(cudaFuncSetAttribute(gx_conv_kernel<C32, KIND>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)std::min<size_t>(100, (smem + 2048) * 100 / (228 * 1024) + 1)));
    gx_conv_kernel<C32, KIND><<<grid, kThreads, smem, s>>>(tmap, k);
    ASRB_CHECK_LAUNCH();
}
}  // namespace

size_t pairbuf_floats(const Plan& P, int ncols) {
    return P.mode == kModeStationary ? (size_t)std::max<int64_t>(P.R, 0) * padded_n(ncols) : 0;
}

void conv(const Plan& P, const ConvArgs& c, cudaStream_t s) {
    ASRB_REQUIRE(P.finished, "gx conv: the plan is not finished");
    if (P.gidx.s != s || P.rare_rs.s != s) const_cast<Plan&>(P).rehome(s);
    const int Cin = c.x.C;
    ASRB_REQUIRE(Cin == 32 || (Cin >= 64 && Cin % 64 == 0), "gx conv: in_channels must be 32 or a multiple of 64");
    const bool c32 = use_c32(Cin, padded_n(c.ncols));
    ASRB_REQUIRE(c.ncols % 8 == 0 && c.ncols >= 8 && c.ncols <= 128,
                 "gx conv: output columns must be a multiple of 8, <= 128 per call (wider banks run as column groups)");
    ASRB_REQUIRE(c.x.rows == P.V_in + 1, "gx conv: the input view must hold the plan's input rows + the zero row");
    // the TMA gather form copies 128 contiguous bytes per row: a plain [32 hi | 32 lo] row only
    const bool tma_ok = !c32 || (Cin == 32 && c.x.hi % 64 == 0 && c.x.lo == c.x.hi + 32);
    ASRB_REQUIRE((c.out.p != nullptr) != (c.out_f32 != nullptr), "gx conv: exactly one of the h2 / fp32 outputs");
    if (P.V == 0) return;
    const int N = padded_n(c.ncols);
    KArgs k{};
    k.V = (int)P.V;
    k.x = c.x.p;
    k.x_pitch = c.x.pitch;
    k.tma_gather = g_tma_gather && tma_ok;
    k.l1_gather = g_l1_gather;
    k.ablate = g_ablate;
    k.trace = 0;
    k.a_hi = c.x.hi;
    k.a_lo = c.x.lo;
    k.chunks = c32 ? Cin / 32 : Cin / 64;
    k.wp = (const uint8_t*)c.wp;
    k.chunk_bytes = (uint32_t)((c32 ? 1 : 2) * N * 128);
    k.slot_bytes = (unsigned long long)k.chunks * k.chunk_bytes;
    k.N = N;
    ASRB_REQUIRE(N <= 128 || !c.imp, "gx conv: the importance variant needs <= 128 output columns");
    k.ncat = (!c32 && N <= 128) ? 1 : 0;  // [B_hi | B_lo] as one operand of width 2 N <= 256
    // accumulator groups (see the kernel): as many as fit next to a second buffer, at most 4
    k.by_slot = 0;
    // two TMEM buffers whenever [main N | corr N] fits twice (the epilogue of a tile then overlaps the next tile's
    // MMAs; for N = 128 that leaves one accumulator group); dev knob gx_single_tmem: round-2a policy (N > 64: one
    // buffer, two groups)
    // accumulator groups as before (stationary: 4 / 4 / 2 / 1 for N = 16 / 32 / 64 / 128; a pair tile has only
    // `chunks` MMA steps), then as many TMEM buffers as fit (<= 4): a buffer is held from the tile's first MMA to
    // the end of its TMEM loads, so more buffers = deeper MMA / epilogue overlap
    k.G = g_acc_groups > 0 ? g_acc_groups : std::max(1, std::min(4, 256 / (2 * N)));
    k.G = std::max(1, std::min(k.G, 512 / (2 * N)));
    k.nbuf = std::max(1, std::min(4, 512 / (2 * N * k.G)));
    if (g_single_tmem && N > 64) {
        k.nbuf = 1;
        k.G = std::max(1, std::min(g_acc_groups > 0 ? g_acc_groups : 4, 512 / (2 * N)));
    }
    k.rare_lp = 4;
    while (k.rare_lp * 4 < N) k.rare_lp <<= 1;
    k.zero_row = (int)P.V_in;
    k.stage_bytes = (c32 ? 1u : 2u) * kATile + k.chunk_bytes;
    k.tx_bytes = k.stage_bytes;
    // shared memory: pipeline stages + 1-2 staging tiles of 128 x (4 N + 16) bytes (rare sums in, results out);
    // the gathers need little depth (2 stages measure the same as 4), so the staging tiles come first
    const size_t budget = 222 * 1024, stage_tile = (size_t)kTM * (4 * N + 16);
    const bool ring = P.mode == kModeStationary && P.R > 0;  // the stationary pass reads the pair buffer through a ring
    k.ring_pairs = (int)std::min<size_t>(32, 8192 / (4 * N));
    k.ring_slots = ring ? (N > 64 ? 3 : 4) : 0;
    const size_t ring_bytes = (size_t)k.ring_slots * k.ring_pairs * N * 4;
    k.nrb = (budget - ring_bytes - 2 * stage_tile) / k.stage_bytes >= 2 ? 2 : 1;
    k.stages = (int)std::max<size_t>(2, std::min<size_t>(8, (budget - ring_bytes - k.nrb * stage_tile) / k.stage_bytes));
    if (g_max_stages > 0) k.stages = std::max(2, std::min(k.stages, g_max_stages));
    const size_t smem = (size_t)k.stages * k.stage_bytes + k.nrb * stage_tile + ring_bytes + 1024;
    k.teams = (!g_one_team && N <= 64 && k.nrb == 2 && k.nbuf >= 2) ? 2 : 1;
    ASRB_REQUIRE(smem <= 227 * 1024 - 4096, "gx conv: shared-memory budget exceeded");  // 4 KB static (barriers, row bounds)
    k.wscale = ldexpf(1.f, -c.scale_exp);
    k.ncols = c.ncols;
    const CUtensorMap tmap = make_row_map(c.x);
    const bool has_rare = P.R > 0;
    char label[96];
    snprintf(label, sizeof(label), "gx_conv K%d %dx%d E%lld", P.K, Cin, c.ncols, (long long)P.E);
    ProfileScope prof(label, s, 2.0 * (double)P.E * Cin * c.ncols);

    auto set_final = [&](KArgs& f) {
        f.bias = c.bias;
        f.norm = c.norm;
        f.relu = c.relu;
        if (c.out.p) {
            ASRB_REQUIRE(c.out.hi % 8 == 0 && c.out.lo % 8 == 0 && c.out.pitch % 8 == 0, "gx conv: output view alignment");
            f.out_h = c.out.p;
            f.out_pitch = c.out.pitch;
            f.out_hi = c.out.hi;
            f.out_lo = c.out.lo;
        } else {
            ASRB_REQUIRE(c.out_f32_pitch % 4 == 0 && c.out_f32_col % 4 == 0, "gx conv: fp32 output alignment");
            f.out_f = c.out_f32 + c.out_f32_col;
            f.out_f_pitch = c.out_f32_pitch;
        }
        if (c.res.p) {
            f.res = c.res.p;
            f.res_pitch = c.res.pitch;
            f.res_hi = c.res.hi;
            f.res_lo = c.res.lo;
        }
    };

    if (has_rare) {
        KArgs r = k;
        r.trace = g_trace_on == 2;
        if (k.chunks == 1) {  // a pair tile is one MMA step: one accumulator group (no change of the result), more buffers
            r.G = 1;
            r.nbuf = std::max(1, std::min(4, 512 / (2 * N)));
        }
        r.gidx = P.pt_gidx.get();
        r.tile_slot = P.pt_slot.get();
        r.out_pos = P.pt_out.get();
        r.num_tiles_dev = P.pt_count.get();
        const int grid = std::min(P.max_pair_tiles, sm_count());
        if (P.mode == kModePairFinal) {
            set_final(r);
            if (c32) launch<true, kKindPairFinal>(tmap, r, grid, smem, s);
            else launch<false, kKindPairFinal>(tmap, r, grid, smem, s);
        } else {
            ASRB_REQUIRE(c.pairbuf != nullptr, "gx conv: pair buffer missing");
            r.pair_out = c.pairbuf;
            if (c32) launch<true, kKindPairBuf>(tmap, r, grid, smem, s);
            else launch<false, kKindPairBuf>(tmap, r, grid, smem, s);
        }
    }
    if (P.mode == kModeStationary) {
        KArgs o = k;
        o.trace = g_trace_on == 1;
        o.gidx = P.gidx.get();
        o.num_tiles = (int)P.T;
        o.D = P.D;
        o.rare_rs = has_rare ? P.rare_rs.get() : nullptr;
        o.row_map = P.d_row_map;
        o.pairbuf = c.pairbuf;
        if (c.imp) {  // one accumulator group per dense slot, weighted in the epilogue
            ASRB_REQUIRE(P.D <= 8 && 2 * N * P.D <= 512, "gx conv: importance variant: too many columns");
            o.imp = c.imp;
            o.rare_in = P.rare_in.get();
            o.by_slot = 1;
            o.G = P.D;
            o.nbuf = 2 * N * P.D * 2 <= 512 ? 2 : 1;
        }
        set_final(o);
        const int grid = (int)std::min<int64_t>(P.T, sm_count());
        if (c32) launch<true, kKindStationary>(tmap, o, grid, smem, s);
        else launch<false, kKindStationary>(tmap, o, grid, smem, s);
    }
}

// ------------------------------------------------------------------------------------------ format helpers
namespace {
// 8 channels per thread
__global__ void __launch_bounds__(256)
from_f32_kernel(const float* __restrict__ x, long long V, int C, int ldx, const float* __restrict__ row_scale,
                const int32_t* __restrict__ rows, __half* __restrict__ out, int pitch, int hi, int lo) {
    const int c8 = C >> 3;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V * c8) return;
    const long long row = i / c8;
    const int col = (int)(i - row * c8) * 8;
    const long long orow = rows ? rows[row] : row;  // input row `row` goes to output row rows[row]
    const float4 a = *reinterpret_cast<const float4*>(x + row * ldx + col);
    const float4 b = *reinterpret_cast<const float4*>(x + row * ldx + col + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (row_scale) {
        const float s = row_scale[row];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= s;
    }
    int overflow = 0;
    split_store8(out + orow * pitch + hi + col, out + orow * pitch + lo + col, v, overflow);
    if (overflow) atomicOr(&g_overflow_flag, 1);
}

__global__ void __launch_bounds__(256)
to_f32_kernel(const __half* __restrict__ x, long long V, int C, int pitch, int hi, int lo, const float* __restrict__ row_scale,
              float* __restrict__ out_f, int ldo, __half* __restrict__ out_h, int opitch, int ohi, int olo) {
    const int c8 = C >> 3;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V * c8) return;
    const long long row = i / c8;
    const int col = (int)(i - row * c8) * 8;
    const uint4 uh = *reinterpret_cast<const uint4*>(x + row * pitch + hi + col);
    const uint4 ul = *reinterpret_cast<const uint4*>(x + row * pitch + lo + col);
    const __half* hh = reinterpret_cast<const __half*>(&uh);
    const __half* ll = reinterpret_cast<const __half*>(&ul);
    float v[8];
    const float s = row_scale ? row_scale[row] : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (__half2float(hh[j]) + __half2float(ll[j])) * s;
    if (out_f) {
        *reinterpret_cast<float4*>(out_f + row * ldo + col) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(out_f + row * ldo + col + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        int overflow = 0;
        split_store8(out_h + row * opitch + ohi + col, out_h + row * opitch + olo + col, v, overflow);
        if (overflow) atomicOr(&g_overflow_flag, 1);
    }
}
}  // namespace

static void zero_last_row(const H2View& v, cudaStream_t s) {
    // only the columns of this view (hi and lo parts); a slice of a shared buffer leaves the rest alone
    ASRB_CUDA(cudaMemsetAsync(v.p + (size_t)(v.rows - 1) * v.pitch + v.hi, 0, (size_t)v.C * 2, s));
    ASRB_CUDA(cudaMemsetAsync(v.p + (size_t)(v.rows - 1) * v.pitch + v.lo, 0, (size_t)v.C * 2, s));
}

void from_f32(const float* x, int64_t V, int C, int ldx, const float* row_scale, const int32_t* rows, H2View out,
              cudaStream_t s) {
    ASRB_REQUIRE(C % 8 == 0 && ldx % 4 == 0 && out.C == C && (rows || out.rows == V + 1), "gx from_f32: bad shapes");
    if (V > 0) {
        from_f32_kernel<<<grid_for((size_t)V * (C / 8), 256), 256, 0, s>>>(x, V, C, ldx, row_scale, rows, out.p, out.pitch,
                                                                          out.hi, out.lo);
        ASRB_CHECK_LAUNCH();
    }
    zero_last_row(out, s);
}

void to_f32(H2View x, int64_t V, float* out, int ldo, cudaStream_t s) {
    ASRB_REQUIRE(x.C % 8 == 0 && ldo % 4 == 0, "gx to_f32: bad shapes");
    if (V == 0) return;
    to_f32_kernel<<<grid_for((size_t)V * (x.C / 8), 256), 256, 0, s>>>(x.p, V, x.C, x.pitch, x.hi, x.lo, nullptr, out, ldo,
                                                                      nullptr, 0, 0, 0);
    ASRB_CHECK_LAUNCH();
}

void scale_rows(H2View x, int64_t V, const float* row_scale, H2View out, cudaStream_t s) {
    ASRB_REQUIRE(x.C % 8 == 0 && out.C == x.C && out.rows == V + 1, "gx scale_rows: bad shapes");
    if (V > 0) {
        to_f32_kernel<<<grid_for((size_t)V * (x.C / 8), 256), 256, 0, s>>>(x.p, V, x.C, x.pitch, x.hi, x.lo, row_scale,
                                                                          nullptr, 0, out.p, out.pitch, out.hi, out.lo);
        ASRB_CHECK_LAUNCH();
    }
    zero_last_row(out, s);
}

void trace_read(unsigned* host, int ctas, cudaStream_t s) {
#ifdef GX_TRACE
    ASRB_CUDA(cudaStreamSynchronize(s));
    ASRB_CUDA(cudaMemcpyFromSymbol(host, g_trace, sizeof(unsigned) * kTraceSlots * std::min(ctas, 256)));
#else
    (void)host, (void)ctas, (void)s;
    throw Error(kInvalidArgument, "gx trace: the library was built without the instrumentation (make TRACE=1)");
#endif
}

int overflow_flag_read_and_clear(cudaStream_t s) {
    int v = 0, z = 0;
    ASRB_CUDA(cudaMemcpyFromSymbolAsync(&v, g_overflow_flag, sizeof(int), 0, cudaMemcpyDeviceToHost, s));
    ASRB_CUDA(cudaMemcpyToSymbolAsync(g_overflow_flag, &z, sizeof(int), 0, cudaMemcpyHostToDevice, s));
    ASRB_CUDA(cudaStreamSynchronize(s));
    return v;
}

}  // namespace gx
}  // namespace asrb
