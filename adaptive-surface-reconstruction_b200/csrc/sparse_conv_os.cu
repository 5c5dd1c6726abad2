// Output-stationary persistent tensor-core sparse convolution.
//
// Used (option sparse_conv_output_stationary) for the convolutions without importance weighting
// (conv2..4 of every block and the decoder blocks) whose filter slice fits the shared-memory ring
// (Cin <= 128, Cout <= 128).  HYBRID form: only the kOsSlots "common" kernel slots (self + the 6
// same-level faces: ~73 % of all pairs, each ~80 % dense) are accumulated output-stationary; the
// sparse finer / coarser slots (the union of ALL slots over 256 rows is ~37 of 55 on adaptive
// clouds, 4.5x the minimal MMA work) stay pair-major and are reduced into the written output by
// the per-tile kernel, followed by the bias / ReLU pass.  The pair-major kernels (sparse_conv_pm.cu,
// sparse_conv_tc.cu) add every 128-pair tile to the output with L2 reductions, and
// measurements on B200 show those reductions — through red.global or through
// cp.reduce.async.bulk alike — to saturate near 2 TB/s, which bounds the small-channel
// levels.  Here nothing is reduced in memory:
//
//   * a persistent CTA owns "super-tiles" of 256 consecutive output rows (2 MMA tiles of
//     128 rows); the grids are Morton ordered, so the rows are spatially compact and the
//     148 CTAs work on neighbouring row ranges at any time (gathers stay L2 resident);
//   * the plan lists, per super-tile, the kernel slots that occur in it ("steps") and, per
//     step, the gather index of every row (-1 = the row has no neighbour in that slot);
//   * per step the filter slice W[slot] (hi | lo) streams through a chunk ring (bulk
//     copies, TMA engine) and is used by both row tiles; the rows' neighbours are gathered
//     by cp.async straight into K-major operand tiles (the tensor core reads the raw fp32
//     bits as tf32 = the "hi" part), converter warps derive the "lo" tiles, one thread issues
//     the 3xTF32 MMAs, and the sum over ALL slots stays in TMEM;
//   * the epilogue adds main + correction accumulators, bias, ReLU and writes every output
//     row exactly once (bulk stores from padded staging rows).  No zero fill, no reduction,
//     no separate epilogue pass.
//
// Warp roles: 0-3 epilogue (TMEM lane quarter = warp), 4-11 converters, 12-13 gather
// loaders, 14 MMA issuer, 15 filter loader.
#include "internal.h"
#include "prims.cuh"
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
using umma::mbar_wait;
constexpr int TM = 128;
constexpr int ST = 256;                 // rows per super-tile
constexpr int KC = umma::kKC;
constexpr int kMaxRaw = 12;
constexpr int kMaxLo = 3;
constexpr int kMaxB = 16;
constexpr int kOsSlots = 7;             // slots < kOsSlots are accumulated in TMEM, the rest pair-major
constexpr int kEpiWarps = 4;
constexpr int kCvtWarps = 8;
constexpr int kLoadWarps = 2;
constexpr int kCvtThreads = kCvtWarps * 32;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kLoadWarp0 = kEpiWarps + kCvtWarps;
constexpr int kMmaWarp = kLoadWarp0 + kLoadWarps;
constexpr int kBWarp = kMmaWarp + 1;
constexpr int kThreads = (kBWarp + 1) * 32;  // 512
}  // namespace

// ------------------------------------------------------------------ plan
// pass 1: per super-tile the slot masks of its two row tiles
__global__ void __launch_bounds__(ST)
os_mask_kernel(const uint8_t* __restrict__ slot, const int64_t* __restrict__ splits, long long V, int limit,
               unsigned long long* __restrict__ mask01, int32_t* __restrict__ count, int* __restrict__ dup) {
    __shared__ unsigned long long s_m[ST / 32];
    const long long v = blockIdx.x * (long long)ST + threadIdx.x;
    unsigned long long m = 0;
    if (v < V) {
        const int64_t e = splits[v + 1];
        for (int64_t j = splits[v]; j < e; ++j) {
            if (slot[j] >= limit) continue;
            const unsigned long long b = 1ULL << slot[j];
            if (m & b) *dup = 1;  // a slot twice in one row: not expressible here, stay pair-major
            m |= b;
        }
    }
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)m);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(m >> 32));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = ((unsigned long long)hi << 32) | lo;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long m0 = s_m[0] | s_m[1] | s_m[2] | s_m[3], m1 = s_m[4] | s_m[5] | s_m[6] | s_m[7];
        // a row tile that has rows but no entry at all would leave its accumulators unwritten
        if (!m0 || (!m1 && blockIdx.x * (long long)ST + TM < V)) *dup = 1;
        mask01[2 * blockIdx.x] = m0;
        mask01[2 * blockIdx.x + 1] = m1;
        count[blockIdx.x] = __popcll(m0 | m1);
    }
}

// pass 2: step metadata (slot | tile flags << 8) and the per-row gather indices
__global__ void __launch_bounds__(ST)
os_fill_kernel(const int32_t* __restrict__ idx, const uint8_t* __restrict__ slot, const int64_t* __restrict__ splits,
               long long V, int limit, const unsigned long long* __restrict__ mask01, const int64_t* __restrict__ off,
               int32_t* __restrict__ meta, int32_t* __restrict__ sidx) {
    const long long v = blockIdx.x * (long long)ST + threadIdx.x;
    const unsigned long long m0 = mask01[2 * blockIdx.x], m1 = mask01[2 * blockIdx.x + 1], m = m0 | m1;
    const long long o = off[blockIdx.x];
    if (threadIdx.x < 64 && ((m >> threadIdx.x) & 1)) {
        const int k = threadIdx.x;
        const int rank = __popcll(m & ((1ULL << k) - 1));
        meta[o + rank] = k | (int)(((m0 >> k) & 1) | (((m1 >> k) & 1) << 1)) << 8;
    }
    if (v < V) {
        const int64_t e = splits[v + 1];
        for (int64_t j = splits[v]; j < e; ++j) {
            const int k = slot[j];
            if (k >= limit) continue;
            const int rank = __popcll(m & ((1ULL << k) - 1));
            sidx[(o + rank) * ST + threadIdx.x] = idx[j];
        }
    }
}

// the entries with slot >= limit as their own CSR (same rows), for the pair-major kernel
__global__ void __launch_bounds__(256)
os_rare_count_kernel(const uint8_t* __restrict__ slot, const int64_t* __restrict__ splits, long long V, int limit,
                     int32_t* __restrict__ count) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= V) return;
    int c = 0;
    const int64_t e = splits[v + 1];
    for (int64_t j = splits[v]; j < e; ++j) c += slot[j] >= limit ? 1 : 0;
    count[v] = c;
}
__global__ void __launch_bounds__(256)
os_rare_fill_kernel(const int32_t* __restrict__ idx, const uint8_t* __restrict__ slot, const int64_t* __restrict__ splits,
                    long long V, int limit, const int64_t* __restrict__ rsplits, int32_t* __restrict__ ridx,
                    uint8_t* __restrict__ rslot) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= V) return;
    int64_t o = rsplits[v];
    const int64_t e = splits[v + 1];
    for (int64_t j = splits[v]; j < e; ++j)
        if (slot[j] >= limit) {
            ridx[o] = idx[j];
            rslot[o] = slot[j];
            ++o;
        }
}

void os_plan_build(ConvPlan& P, const int32_t* d_idx, const uint8_t* d_slot, const int64_t* d_splits, int64_t V_out,
                   int K, cudaStream_t s) {
    P.os_ok = false;
    if (K > 64 || V_out <= 0) return;
    const long long NU = (V_out + ST - 1) / ST;
    DevBuf<unsigned long long> mask01((size_t)NU * 2, s);
    DevBuf<int32_t> count((size_t)NU, s);
    DevBuf<int> dup(1, s);
    ASRB_CUDA(cudaMemsetAsync(dup.get(), 0, sizeof(int), s));
    const int limit = kOsSlots;
    os_mask_kernel<<<(unsigned)NU, ST, 0, s>>>(d_slot, d_splits, V_out, limit, mask01.get(), count.get(), dup.get());
    ASRB_CHECK_LAUNCH();
    P.os_off.alloc((size_t)NU + 1, s);
    exclusive_sum_i32_to_i64(count.get(), P.os_off.get(), (size_t)NU, s);
    const int64_t NS = d2h_scalar(P.os_off.get() + NU, s);
    if (d2h_scalar(dup.get(), s)) return;
    P.os_steps = NS;
    P.os_tiles = NU;
    P.os_meta.alloc((size_t)NS, s);
    P.os_idx.alloc((size_t)NS * ST, s);
    ASRB_CUDA(cudaMemsetAsync(P.os_idx.get(), 0xff, (size_t)NS * ST * sizeof(int32_t), s));
    os_fill_kernel<<<(unsigned)NU, ST, 0, s>>>(d_idx, d_slot, d_splits, V_out, limit, mask01.get(), P.os_off.get(),
                                               P.os_meta.get(), P.os_idx.get());
    ASRB_CHECK_LAUNCH();
    // the remaining (finer / coarser) entries: their own pair-major plan
    DevBuf<int32_t> rcount((size_t)V_out, s);
    DevBuf<int64_t> rsplits((size_t)V_out + 1, s);
    os_rare_count_kernel<<<grid_for(V_out, 256), 256, 0, s>>>(d_slot, d_splits, V_out, limit, rcount.get());
    ASRB_CHECK_LAUNCH();
    exclusive_sum_i32_to_i64(rcount.get(), rsplits.get(), (size_t)V_out, s);
    const int64_t E_rare = d2h_scalar(rsplits.get() + V_out, s);
    P.rare.reset();
    P.os_pairs = P.E - E_rare;
    if (E_rare > 0) {
        DevBuf<int32_t> ridx((size_t)E_rare, s);
        DevBuf<uint8_t> rslot((size_t)E_rare, s);
        os_rare_fill_kernel<<<grid_for(V_out, 256), 256, 0, s>>>(d_idx, d_slot, d_splits, V_out, limit, rsplits.get(),
                                                                 ridx.get(), rslot.get());
        ASRB_CHECK_LAUNCH();
        P.rare = std::make_unique<ConvPlan>();
        conv_plan_build(*P.rare, ridx.get(), rslot.get(), rsplits.get(), V_out, E_rare, K, s, false);
    }
    P.os_ok = true;
}

// ------------------------------------------------------------------ kernel
struct OsArgs {
    const float* x;
    const float* wp;          // packed filters [slot][chunk][hi|lo][n_pad x KC]
    const int64_t* os_off;    // [NU + 1]
    const int32_t* os_meta;   // [NS] slot | flags << 8
    const int32_t* os_idx;    // [NS][256]
    const float* bias;
    float* out;
    long long V, NU;
    int Cin, Cout, relu;
    int nc, nch;              // output columns (= Cout), 16-channel chunks
    int R, Q, NBS, UB;        // ring depths; UB = TMEM super-tile buffers (1 or 2)
    int eb;                   // epilogue column block
    uint32_t off_raw, off_lo, off_t, off_bias;
    int dbg;
};

// dev instrumentation (option pm_debug = 1), CTA 0, cycles: [0] loader total [1] wait raw_empty
// [2] cvt total [3] wait raw_full [4] wait lo_empty [5] mma total [6] wait b_full [7] wait lo_full
// [8] wait acc_empty [9] filter total [10] wait b_empty [11] epi total [12] wait acc_full
// [13] stages [14] steps [15] super-tiles
__device__ unsigned long long g_os_dbg[16];
#define OS_T0() const long long t0__ = a.dbg ? clock64() : 0
#define OS_ACC(i) do { if (a.dbg) dbg_w[i] += (unsigned long long)(clock64() - t0__); } while (0)
#define OS_FLUSH(i) do { if (a.dbg && blockIdx.x == 0) g_os_dbg[i] += dbg_w[i]; } while (0)

// step metadata of one super-tile, spread over the lanes of a warp (<= 64 steps)
struct StepList {
    int m0, m1, n;
    __device__ __forceinline__ int get(int i) const { return __shfl_sync(0xffffffffu, i < 32 ? m0 : m1, i & 31); }
};
__device__ __forceinline__ StepList load_steps(const OsArgs& a, long long u, int lane, long long& j0) {
    StepList L;
    L.m0 = L.m1 = 0;
    L.n = 0;
    j0 = 0;
    if (u < a.NU) {
        j0 = __ldg(a.os_off + u);
        L.n = (int)(__ldg(a.os_off + u + 1) - j0);
        if (lane < L.n) L.m0 = __ldg(a.os_meta + j0 + lane);
        if (lane + 32 < L.n) L.m1 = __ldg(a.os_meta + j0 + 32 + lane);
    }
    return L;
}

__global__ void __launch_bounds__(kThreads, 1)
sparse_conv_os_kernel(OsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t raw_full[kMaxRaw];
    __shared__ uint64_t raw_empty[kMaxRaw];
    __shared__ uint64_t lo_full[kMaxLo];
    __shared__ uint64_t lo_empty[kMaxLo];
    __shared__ uint64_t b_full[kMaxB];
    __shared__ uint64_t b_empty[kMaxB];
    __shared__ uint64_t acc_full[2];
    __shared__ uint64_t acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.R, Q = a.Q, NBS = a.NBS, UB = a.UB;
    const int nc = a.nc, nch = a.nch;
    const uint32_t need = (uint32_t)UB * 2 * 2 * nc;  // buffers x tiles x (main + correction)
    const uint32_t ncols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    const uint32_t b_chunk_bytes = 2 * (uint32_t)nc * KC * 4;

    if (warp == kMmaWarp) {
        umma::tmem_alloc(&tmem_slot, ncols);
        if (lane == 0) {
            for (int i = 0; i < R; ++i) {
                umma::mbar_init(&raw_full[i], kLoadThreads);
                umma::mbar_init(&raw_empty[i], 1);
            }
            for (int i = 0; i < Q; ++i) {
                umma::mbar_init(&lo_full[i], kCvtThreads);
                umma::mbar_init(&lo_empty[i], 1);
            }
            for (int i = 0; i < NBS; ++i) {
                umma::mbar_init(&b_full[i], 1);
                umma::mbar_init(&b_empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                umma::mbar_init(&acc_full[i], 1);
                umma::mbar_init(&acc_empty[i], kEpiWarps * 32);
            }
            umma::fence_barrier_init();
        }
    }
    float* s_bias = reinterpret_cast<float*>(smem + a.off_bias);
    for (int i = tid; i < nc; i += kThreads) s_bias[i] = a.bias ? a.bias[i] : 0.f;
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint8_t* sB = smem;
    uint8_t* sRaw = smem + a.off_raw;
    uint8_t* sLo = smem + a.off_lo;
    const long long u0 = blockIdx.x, du = gridDim.x;
    unsigned long long dbg_w[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_role0 = a.dbg ? clock64() : 0;

    if (warp < kEpiWarps) {
        // ================================================================ epilogue
        const int row = tid;
        float* T = reinterpret_cast<float*>(smem + a.off_t) + (size_t)row * (a.eb + 4);
        int n = 0;
        for (long long u = u0; u < a.NU; u += du, ++n) {
            const int ub = n % UB, use = n / UB;
            {
                OS_T0();
                mbar_wait(&acc_full[ub], use & 1);
                OS_ACC(12);
            }
            umma::tc_fence_after();
            for (int t = 0; t < 2; ++t) {
                const long long orow = u * ST + t * TM + row;
                const uint32_t t_acc = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ub * 2 + t) * 2 * nc;
                for (int nb = 0; nb < nc; nb += a.eb) {
                    umma::bulk_wait_read();
                    for (int n0 = 0; n0 < a.eb; n0 += 32) {
                        uint32_t m[32], c[32];
                        umma::tmem_ld32_issue(t_acc + nb + n0, m);
                        umma::tmem_ld32_issue(t_acc + nc + nb + n0, c);
                        umma::tmem_ld_wait32(m);
                        umma::tmem_ld_wait32(c);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(s_bias + nb + n0 + j);
                            float4 e;
                            e.x = __uint_as_float(m[j]) + __uint_as_float(c[j]) + b.x;
                            e.y = __uint_as_float(m[j + 1]) + __uint_as_float(c[j + 1]) + b.y;
                            e.z = __uint_as_float(m[j + 2]) + __uint_as_float(c[j + 2]) + b.z;
                            e.w = __uint_as_float(m[j + 3]) + __uint_as_float(c[j + 3]) + b.w;
                            if (a.relu) {
                                e.x = fmaxf(e.x, 0.f); e.y = fmaxf(e.y, 0.f);
                                e.z = fmaxf(e.z, 0.f); e.w = fmaxf(e.w, 0.f);
                            }
                            *reinterpret_cast<float4*>(T + n0 + j) = e;
                        }
                    }
                    if (t == 1 && nb + a.eb >= nc) {  // both tiles are out of TMEM
                        umma::tc_fence_before();
                        mbar_arrive(&acc_empty[ub]);
                    }
                    umma::fence_proxy_async();
                    if (orow < a.V) umma::bulk_store(a.out + (size_t)orow * a.Cout + nb, T, (uint32_t)a.eb * 4);
                    umma::bulk_commit();
                }
            }
        }
        umma::bulk_wait_all();
        if (a.dbg && blockIdx.x == 0 && tid == 0) {
            g_os_dbg[11] += (unsigned long long)(clock64() - t_role0);
            OS_FLUSH(12);
        }
    } else if (warp < kLoadWarp0) {
        // ================================================================ converters (raw -> lo)
        const int ct = tid - kEpiWarps * 32;
        const int kq = ct & 3, rsub = ct >> 2;
        const uint32_t off0 = (uint32_t)(rsub >> 3) * kA_SBO + (uint32_t)kq * kA_LBO + (uint32_t)(rsub & 7) * 16;
        const uint32_t off1 = off0 + 8 * kA_SBO;
        int r = 0, rph = 0, q = 0, qph = 0, done = 0;
        for (long long u = u0; u < a.NU; u += du) {
            long long j0;
            const StepList L = load_steps(a, u, lane, j0);
            int stages = 0;
            for (int i = 0; i < L.n; ++i) stages += __popc((L.get(i) >> 8) & 3);
            stages *= nch;
            for (int sidx = 0; sidx < stages; ++sidx, ++done) {
                {
                    OS_T0();
                    mbar_wait(&raw_full[r], rph);
                    OS_ACC(3);
                }
                const uint8_t* src = sRaw + (size_t)r * kATileBytes;
                const float4 x0 = *reinterpret_cast<const float4*>(src + off0);
                const float4 x1 = *reinterpret_cast<const float4*>(src + off1);
                float4 l0, l1;
                l0.x = x0.x - umma::tf32_hi(x0.x); l0.y = x0.y - umma::tf32_hi(x0.y);
                l0.z = x0.z - umma::tf32_hi(x0.z); l0.w = x0.w - umma::tf32_hi(x0.w);
                l1.x = x1.x - umma::tf32_hi(x1.x); l1.y = x1.y - umma::tf32_hi(x1.y);
                l1.z = x1.z - umma::tf32_hi(x1.z); l1.w = x1.w - umma::tf32_hi(x1.w);
                if (done >= Q) {
                    OS_T0();
                    mbar_wait(&lo_empty[q], qph ^ 1);
                    OS_ACC(4);
                }
                uint8_t* dst = sLo + (size_t)q * kATileBytes;
                *reinterpret_cast<float4*>(dst + off0) = l0;
                *reinterpret_cast<float4*>(dst + off1) = l1;
                umma::fence_proxy_async();
                mbar_arrive(&lo_full[q]);
                if (++r == R) {
                    r = 0;
                    rph ^= 1;
                }
                if (++q == Q) {
                    q = 0;
                    qph ^= 1;
                }
            }
        }
        if (a.dbg && blockIdx.x == 0 && ct == 0) {
            g_os_dbg[2] += (unsigned long long)(clock64() - t_role0);
            OS_FLUSH(3);
            OS_FLUSH(4);
        }
    } else if (warp < kMmaWarp) {
        // ================================================================ gather loaders
        const int lw = warp - kLoadWarp0;
        const int kq = lane & 3, rl = lane >> 2;
        const uint32_t off_base = (uint32_t)(lw * 8) * kA_SBO + (uint32_t)kq * kA_LBO + (uint32_t)rl * 16;
        int r = 0, rph = 0, done = 0;
        for (long long u = u0; u < a.NU; u += du) {
            long long j0;
            const StepList L = load_steps(a, u, lane, j0);
            // units = (step i, row tile t) that have rows; the gather indices of the next unit are
            // fetched while the current one is being issued
            auto next_unit = [&](int& i, int& t) {
                for (;;) {
                    if (++t == 2) {
                        t = 0;
                        ++i;
                    }
                    if (i >= L.n) return false;
                    if ((L.get(i) >> (8 + t)) & 1) return true;
                }
            };
            auto load_pins = [&](int i, int t, int (&pin)[8]) {
                const int32_t* ip = a.os_idx + (j0 + i) * ST + t * TM + lw * 64 + rl;
#pragma unroll
                for (int j = 0; j < 8; ++j) pin[j] = __ldg(ip + 8 * j);
            };
            int i = 0, t = -1;
            bool have = next_unit(i, t);
            int pin[8], pin_next[8];
            if (have) load_pins(i, t, pin);
            while (have) {
                int ni = i, nt = t;
                const bool have_next = next_unit(ni, nt);
                if (have_next) load_pins(ni, nt, pin_next);
                for (int c = 0; c < nch; ++c, ++done) {
                    if (done >= R) {
                        OS_T0();
                        mbar_wait(&raw_empty[r], rph ^ 1);
                        OS_ACC(1);
                    }
                    const uint32_t dst = umma::smem_u32(sRaw + (size_t)r * kATileBytes) + off_base;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const bool ok = pin[j] >= 0;
                        const float* src = ok ? a.x + (size_t)pin[j] * a.Cin + c * KC + kq * 4 : a.x;
                        umma::cp_async16_cg(dst + (uint32_t)j * kA_SBO, src, ok ? 16u : 0u);
                    }
                    umma::cp_async_arrive_noinc(&raw_full[r]);
                    if (++r == R) {
                        r = 0;
                        rph ^= 1;
                    }
                }
                i = ni;
                t = nt;
                have = have_next;
#pragma unroll
                for (int j = 0; j < 8; ++j) pin[j] = pin_next[j];
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (a.dbg && blockIdx.x == 0 && tid == kLoadWarp0 * 32) {
            g_os_dbg[0] += (unsigned long long)(clock64() - t_role0);
            OS_FLUSH(1);
        }
    } else if (warp == kMmaWarp) {
        // ================================================================ MMA issuer (one thread: every extra
        // instruction on its path costs issue rate; the step words are read
        // straight from global memory, one step ahead)
        if (lane == 0) {
            const uint32_t idesc = umma::make_idesc_tf32(128, nc), idesc2 = umma::make_idesc_tf32(128, 2 * nc);
            const uint64_t d_raw0 = umma::desc_base(kA_LBO, kA_SBO) + (umma::smem_u32(sRaw) >> 4);
            const uint64_t d_lo0 = umma::desc_base(kA_LBO, kA_SBO) + (umma::smem_u32(sLo) >> 4);
            const uint64_t d_b0 = umma::desc_base(kB_LBO, kB_SBO) + (umma::smem_u32(sB) >> 4);
            int r = 0, q = 0, qph = 0, bi = 0, n = 0;
            for (long long u = u0; u < a.NU; u += du, ++n) {
                const long long j0 = __ldg(a.os_off + u);
                const int nsteps = (int)(__ldg(a.os_off + u + 1) - j0);
                const int ub = n % UB, use = n / UB;
                if (use > 0) {
                    OS_T0();
                    mbar_wait(&acc_empty[ub], (use - 1) & 1);
                    OS_ACC(8);
                }
                umma::tc_fence_after();
                dbg_w[15] += 1;
                dbg_w[14] += nsteps;
                bool started0 = false, started1 = false;
                int meta_next = nsteps > 0 ? __ldg(a.os_meta + j0) : 0;
                for (int i = 0; i < nsteps; ++i, bi += nch) {
                    const int flags = (meta_next >> 8) & 3;
                    if (i + 1 < nsteps) meta_next = __ldg(a.os_meta + j0 + i + 1);
                    const int last_t = (flags & 2) ? 1 : 0;
                    bool first_tile = true;
                    for (int t = 0; t < 2; ++t) {
                        if (!((flags >> t) & 1)) continue;
                        const uint32_t t_main = tmem + (uint32_t)(ub * 2 + t) * 2 * nc, t_corr = t_main + nc;
                        const bool started = t ? started1 : started0;
                        for (int c = 0; c < nch; ++c) {
                            const int bs = (bi + c) % NBS;
                            if (first_tile) {
                                OS_T0();
                                mbar_wait(&b_full[bs], ((bi + c) / NBS) & 1);
                                OS_ACC(6);
                            }
                            {
                                OS_T0();
                                mbar_wait(&lo_full[q], qph);
                                OS_ACC(7);
                            }
                            dbg_w[13] += 1;
                            umma::tc_fence_after();
                            const uint64_t a_hi0 = d_raw0 + (uint64_t)(r * (int)(kATileBytes >> 4));
                            const uint64_t a_lo0 = d_lo0 + (uint64_t)(q * (int)(kATileBytes >> 4));
                            const uint64_t b_hi0 = d_b0 + (uint64_t)(bs * (int)(b_chunk_bytes >> 4));
                            const bool acc0 = started || c > 0;
#pragma unroll
                            for (int ks = 0; ks < KC / 8; ++ks) {
                                const uint64_t oa = (uint64_t)(ks * 2 * (kA_LBO >> 4)), ob = (uint64_t)(ks * 2 * (kB_LBO >> 4));
                                // A_hi x [B_hi | B_lo] -> (main | correction) in one MMA of width 2 nc (sparse_conv_tc.cu)
                                umma::mma_tf32(t_main, a_hi0 + oa, b_hi0 + ob, idesc2, acc0 || ks > 0);
                                umma::mma_tf32_acc(t_corr, a_lo0 + oa, b_hi0 + ob, idesc);
                            }
                            umma::mma_commit(&raw_empty[r]);
                            umma::mma_commit(&lo_empty[q]);
                            if (t == last_t) umma::mma_commit(&b_empty[bs]);
                            if (++r == R) r = 0;
                            if (++q == Q) {
                                q = 0;
                                qph ^= 1;
                            }
                        }
                        if (t) started1 = true;
                        else started0 = true;
                        first_tile = false;
                    }
                }
                umma::mma_commit(&acc_full[ub]);
            }
            if (a.dbg && blockIdx.x == 0) {
                g_os_dbg[5] += (unsigned long long)(clock64() - t_role0);
                OS_FLUSH(6);
                OS_FLUSH(7);
                OS_FLUSH(8);
                OS_FLUSH(13);
                OS_FLUSH(14);
                OS_FLUSH(15);
            }
        }
    } else {
        // ================================================================ filter loader (lane 0; the warp loops together)
        const int chunks_total = nch;
        int bi = 0;
        for (long long u = u0; u < a.NU; u += du) {
            long long j0;
            const StepList L = load_steps(a, u, lane, j0);
            for (int i = 0; i < L.n; ++i) {
                const int slot = L.get(i) & 0xff;
                const float* src = a.wp + (size_t)slot * chunks_total * 2 * nc * KC;
                for (int c = 0; c < nch; ++c, ++bi) {
                    const int bs = bi % NBS;
                    if (bi >= NBS) {
                        OS_T0();
                        mbar_wait(&b_empty[bs], ((bi / NBS) - 1) & 1);
                        OS_ACC(10);
                    }
                    if (lane == 0) {
                        umma::mbar_arrive_expect_tx(&b_full[bs], b_chunk_bytes);
                        bulk_copy_g2s(sB + (size_t)bs * b_chunk_bytes, src + (size_t)c * 2 * nc * KC, b_chunk_bytes, &b_full[bs]);
                    }
                    __syncwarp();
                }
            }
        }
        if (a.dbg && blockIdx.x == 0 && lane == 0) {
            g_os_dbg[9] += (unsigned long long)(clock64() - t_role0);
            OS_FLUSH(10);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) umma::tmem_dealloc(tmem, ncols);
}

static bool g_os_enabled = false;  // see DESIGN.md: the slot union of 256 rows is ~37 of 55 slots on adaptive clouds
static int g_os_debug = 0;
void sparse_conv_os_debug(int on) { g_os_debug = on; }
void sparse_conv_os_enable(bool on) { g_os_enabled = on; }
bool sparse_conv_os_enabled() { return g_os_enabled; }

bool sparse_conv_os_supported(const ConvPlan& P, int Cin, int Cout) {
    return g_os_enabled && P.os_ok && Cin % KC == 0 && Cin <= 128 && Cout % 32 == 0 && Cout <= 128;
}

void sparse_conv_os(const ConvPlan& P, const float* x, const float* wp, int Cin, int Cout, const float* bias, int relu,
                    float* out, cudaStream_t s) {
    OsArgs a;
    a.x = x;
    a.wp = wp;
    a.os_off = P.os_off.get();
    a.os_meta = P.os_meta.get();
    a.os_idx = P.os_idx.get();
    a.bias = bias;
    a.out = out;
    a.V = P.V_out;
    a.NU = P.os_tiles;
    a.Cin = Cin;
    a.Cout = Cout;
    a.relu = relu;
    a.nc = Cout;
    a.nch = Cin / KC;
    a.UB = Cout <= 64 ? 2 : 1;
    const size_t budget = 225 * 1024;
    const size_t b_chunk = (size_t)2 * a.nc * KC * 4;
    a.NBS = (2 * a.nch * b_chunk <= 64 * 1024) ? 2 * a.nch : a.nch;
    const size_t b_bytes = (size_t)a.NBS * b_chunk;
    a.Q = b_bytes > 64 * 1024 ? 2 : kMaxLo;
    a.eb = std::min(a.nc, 64);
    auto raw_stages = [&]() {
        const size_t t_bytes = (size_t)TM * (a.eb + 4) * sizeof(float);
        const size_t fixed = b_bytes + (size_t)a.Q * kATileBytes + t_bytes + (size_t)a.nc * 4;
        return fixed >= budget ? 0 : (int)std::min<size_t>(kMaxRaw, (budget - fixed) / kATileBytes);
    };
    a.R = raw_stages();
    if (a.R < 7 && a.eb > 32) {
        a.eb = 32;
        a.R = raw_stages();
    }
    ASRB_REQUIRE(a.R >= 3, "sparse_conv_os: shared-memory budget");
    const size_t t_bytes = (size_t)TM * (a.eb + 4) * sizeof(float);
    a.off_raw = (uint32_t)b_bytes;
    a.off_lo = (uint32_t)(b_bytes + (size_t)a.R * kATileBytes);
    a.off_t = (uint32_t)(a.off_lo + (size_t)a.Q * kATileBytes);
    a.off_bias = (uint32_t)(a.off_t + t_bytes);
    const size_t smem = a.off_bias + (size_t)a.nc * 4;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        ASRB_CUDA(cudaGetDevice(&dev));
        ASRB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    ASRB_CUDA(cudaFuncSetAttribute(sparse_conv_os_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    char label[96];
    snprintf(label, sizeof(label), "sparse_conv_tile/os K%d %dx%d E%lld", P.K, Cin, Cout, (long long)P.os_pairs);
    ProfileScope prof(label, s, 2.0 * (double)P.os_pairs * Cin * Cout);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(n_sm, P.os_tiles));
    a.dbg = g_os_debug == 1;
    if (a.dbg) {
        unsigned long long z[16] = {0};
        ASRB_CUDA(cudaMemcpyToSymbol(g_os_dbg, z, sizeof(z)));
    }
    sparse_conv_os_kernel<<<grid, kThreads, smem, s>>>(a);
    ASRB_CHECK_LAUNCH();
    if (a.dbg) {
        unsigned long long h[16];
        ASRB_CUDA(cudaStreamSynchronize(s));
        ASRB_CUDA(cudaMemcpyFromSymbol(h, g_os_dbg, sizeof(h)));
        fprintf(stderr, "[os] %s R %d Q %d NBS %d UB %d eb %d | supertiles %llu steps %llu stages %llu | load %llu "
                        "wait_raw_empty %llu | cvt %llu wait_raw_full %llu wait_lo_empty %llu | mma %llu wait_b %llu "
                        "wait_lo_full %llu wait_acc_empty %llu | filt %llu wait_b_empty %llu | epi %llu wait_acc_full %llu\n",
                label, a.R, a.Q, a.NBS, a.UB, a.eb, h[15], h[14], h[13], h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7],
                h[8], h[9], h[10], h[11], h[12]);
    }
}

}  // namespace asrb
