// Persistent pair-major tensor-core sparse convolution (the default contraction kernel).
//
// Same mathematics and plan as sparse_conv_tc.cu (slot-sorted pairs inside row blocks,
// 128-pair tiles, tcgen05.mma kind::tf32 with the 3xTF32 split, main and correction
// accumulators kept apart), but organised the way the B200 wants it:
//
//   * ONE persistent CTA per SM walks a contiguous range of work items.  An item is
//     (group = (row block, slot), part = (128-channel slice of Cin, <=128-column slice of
//     Cout), 128-pair tile); items are ordered group-major, then part, then tile, so
//     consecutive items share their filter part.
//   * The filter part W[slot][k0:k0+kc, n0:n0+nc] (hi and lo, <= 128 KB) is RESIDENT in
//     shared memory: loaded once with bulk copies (TMA engine) when the (slot, part) of the
//     range changes, not once per tile.  ncu on the per-tile kernel showed the filter chunks
//     to be 2/3 of its L2->SM traffic.
//   * 8 producer warps gather the pairs' input rows (16-byte coalesced loads, 8 chunks per
//     thread in flight across tile boundaries), split them into tf32 hi / lo and stage them
//     in a ring of K-major operand tiles (bank-conflict-free layout, umma.cuh).
//   * One thread issues the MMAs into one of TWO TMEM accumulator sets, so the epilogue of
//     tile i overlaps the MMAs of tile i + 1.
//   * 4 epilogue warps (thread = pair = TMEM lane) read the accumulators, apply the
//     per-pair importance, stage the pair's output segment in shared memory and add it to
//     the output row with ONE bulk reduction (cp.reduce.async.bulk.add.f32) per pair and
//     column block: the scatter runs on the TMA engine instead of the LSU pipe.
//
// Splitting Cin into 128-channel parts also keeps every accumulation chain at <= 16
// tensor-pipe accumulations before the exact fp32 reduction in L2 (accuracy note in
// sparse_conv_tc.cu).
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
using umma::mbar_wait;
constexpr int TM = 128;
constexpr int KC = umma::kKC;
constexpr int kMaxRaw = 12;             // raw (= hi operand) ring
constexpr int kMaxLo = 3;               // lo operand ring
constexpr int kEpiWarps = 8;            // warps 0..7: TMEM lane quarter = warp & 3, column half = warp >> 2
constexpr int kCvtWarps = 4;            // warps 8..11
constexpr int kLoadWarps = 2;           // warps 12..13
constexpr int kCvtThreads = kCvtWarps * 32;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kLoadWarp0 = kEpiWarps + kCvtWarps;
constexpr int kMmaWarp = kLoadWarp0 + kLoadWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 480
}  // namespace

struct PmArgs {
    const float* x;
    const float* wp;  // packed filters [slot][chunk][hi|lo][n_pad x KC]
    const int32_t* p_in;
    const int32_t* p_out;
    const uint32_t* perm;
    const long long* g_begin;  // [G + 1]
    const int* g_tile0;        // [G + 1]
    const float* imp_in;
    const float* imp_entry;
    float* out;
    int G, K, num_blocks;  // groups in sorted order: (block, slot 0) for all blocks, then per block slots 1..K-1
    int Cin, Cout, n_pad, imp_col;
    int nparts;   // column parts; part index p = kpart * nparts + npart
    int parts;    // kparts * nparts
    int kc, nc;   // input channels / output columns per part
    int nch;      // kc / 16
    int R, Q;     // depth of the raw and lo rings
    int eb;       // epilogue column block (<= nc, multiple of 32)
    uint32_t off_raw, off_lo, off_t;  // byte offsets of the rings and the epilogue staging rows
    int dbg;      // dev: block 0 records the cycles each role spends waiting (g_pm_dbg)
};

// dev instrumentation (option pm_debug): [0] converter total, [1] wait raw_full, [2] mma total,
// [3] wait lo_full, [4] wait acc_empty, [5] wait b, [6] epilogue total, [7] wait acc_full,
// [8] wait staging read, [9] items, [10] loader total, [11] loader wait raw_empty, [12] wait lo_empty
__device__ unsigned long long g_pm_dbg[16];
#define PM_T0() const long long t0__ = a.dbg ? clock64() : 0
#define PM_ACC(i) do { if (a.dbg) dbg_w[i] += (unsigned long long)(clock64() - t0__); } while (0)
#define PM_FLUSH(i) do { if (a.dbg && blockIdx.x == 0) g_pm_dbg[i] += dbg_w[i]; } while (0)

struct ItemIter {
    int g, p, t, nt;
    long long base, len;
};

__device__ __forceinline__ void iter_load_group(const PmArgs& a, ItemIter& it) {
    while (it.g < a.G) {
        const long long b = __ldg(a.g_begin + it.g), e = __ldg(a.g_begin + it.g + 1);
        if (e > b) {
            it.base = b;
            it.len = e - b;
            it.nt = (int)((e - b + TM - 1) / TM);
            return;
        }
        ++it.g;
    }
    it.nt = 0;
    it.len = 0;
    it.base = 0;
}
__device__ __forceinline__ void iter_next(const PmArgs& a, ItemIter& it) {
    if (++it.t >= it.nt) {
        it.t = 0;
        if (++it.p >= a.parts) {
            it.p = 0;
            ++it.g;
            iter_load_group(a, it);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
sparse_conv_pm_kernel(PmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t raw_full[kMaxRaw];   // loaders' cp.async have landed
    __shared__ uint64_t raw_empty[kMaxRaw];  // the MMAs that read the raw tile are done
    __shared__ uint64_t lo_full[kMaxLo];     // converters wrote the lo tile
    __shared__ uint64_t lo_empty[kMaxLo];
    __shared__ uint64_t b_full;
    __shared__ uint64_t acc_full[2];
    __shared__ uint64_t acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_start[4];  // g, p, t, n_items

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.R, Q = a.Q;
    const int nc = a.nc, nch = a.nch;
    const uint32_t need = 4 * (uint32_t)nc;  // 2 buffers x (main + correction)
    const uint32_t ncols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    const uint32_t b_chunk_bytes = 2 * (uint32_t)nc * KC * 4;  // hi + lo of one 16-channel chunk

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
            umma::mbar_init(&b_full, 1);
            for (int i = 0; i < 2; ++i) {
                umma::mbar_init(&acc_full[i], 1);
                umma::mbar_init(&acc_empty[i], kEpiWarps * 32);
            }
            umma::fence_barrier_init();
        }
    } else if (tid == 0) {
        // this CTA's item range -> starting (group, part, tile)
        const long long total = (long long)__ldg(a.g_tile0 + a.G) * a.parts;
        const long long i0 = total * blockIdx.x / gridDim.x, i1 = total * (blockIdx.x + 1) / gridDim.x;
        int lo = 0, hi = a.G;  // last g with parts * g_tile0[g] <= i0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)__ldg(a.g_tile0 + mid) * a.parts <= i0) lo = mid;
            else hi = mid - 1;
        }
        int p = 0, t = 0;
        if (i1 > i0) {
            const int nt = __ldg(a.g_tile0 + lo + 1) - __ldg(a.g_tile0 + lo);
            const long long r = i0 - (long long)__ldg(a.g_tile0 + lo) * a.parts;
            p = (int)(r / nt);
            t = (int)(r % nt);
        }
        s_start[0] = lo;
        s_start[1] = p;
        s_start[2] = t;
        s_start[3] = (int)(i1 - i0);
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_items = s_start[3];
    const int total_steps = n_items * nch;
    ItemIter it;
    it.g = s_start[0];
    it.p = s_start[1];
    it.t = s_start[2];
    it.nt = 0;
    it.base = it.len = 0;
    if (n_items > 0) iter_load_group(a, it);

    uint8_t* sB = smem;
    uint8_t* sRaw = smem + a.off_raw;
    uint8_t* sLo = smem + a.off_lo;
    unsigned long long dbg_w[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // registers; dead unless a.dbg

    if (warp < kEpiWarps) {
        // ================================================================ epilogue
        // 8 warps: the accumulator read-out (not the MMAs) was the busiest role with 4.  A thread
        // owns one pair (TMEM lane) and one half of the part's columns.
        const int row = (warp & 3) * 32 + lane;  // pair inside the tile = TMEM lane
        const int half = warp >> 2;
        const int hc = nc >> 1;                  // columns per thread
        float* T = reinterpret_cast<float*>(smem + a.off_t) + (size_t)tid * (a.eb + 4);
        auto meta = [&](const ItemIter& q, int& o, float& imp) {
            const int cnt = (int)min((long long)TM, q.len - (long long)q.t * TM);
            const long long s0 = q.base + (long long)q.t * TM;
            o = -1;
            imp = 1.f;
            if (row < cnt) {
                o = __ldg(a.p_out + s0 + row);
                if (a.imp_in) imp = __ldg(a.imp_in + __ldg(a.p_in + s0 + row));
                if (a.imp_entry) imp *= __ldg(a.imp_entry + __ldg(a.perm + s0 + row));
            }
        };
        int o_next = -1;
        float imp_next = 1.f;
        if (n_items > 0) meta(it, o_next, imp_next);
        const long long te0 = a.dbg ? clock64() : 0;
        for (int n = 0; n < n_items; ++n) {
            const int buf = n & 1, use = n >> 1;
            const int o = o_next;
            const float imp = imp_next;
            const int col0 = (it.p % a.nparts) * nc + half * hc;  // first column of this thread
            iter_next(a, it);
            const int ncols_valid = min(hc, a.Cout - col0);
            {
                PM_T0();
                mbar_wait(&acc_full[buf], use & 1);
                PM_ACC(7);
            }
            umma::tc_fence_after();
            const uint32_t t_acc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * 2 * nc + half * hc;
            for (int nb = 0; nb < hc; nb += a.eb) {
                {
                    PM_T0();
                    umma::bulk_wait_read();  // the staging row is free again
                    PM_ACC(8);
                }
                for (int n0 = 0; n0 < a.eb; n0 += 16) {
                    uint32_t m[16], c[16];
                    {
                        PM_T0();
                        umma::tmem_ld16_issue(t_acc + nb + n0, m);
                        umma::tmem_ld16_issue(t_acc + nc + nb + n0, c);
                        umma::tmem_ld_wait16(m);
                        umma::tmem_ld_wait16(c);
                        PM_ACC(13);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float e[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            e[q] = __uint_as_float(m[j + q]) + __uint_as_float(c[j + q]);
                            if (col0 + nb + n0 + j + q >= a.imp_col) e[q] *= imp;
                        }
                        *reinterpret_cast<float4*>(T + n0 + j) = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
                if (nb + a.eb >= hc) {  // the accumulators are in registers / staged: release the TMEM buffer
                    umma::tc_fence_before();
                    mbar_arrive(&acc_empty[buf]);
                }
                {
                    PM_T0();
                    umma::fence_proxy_async();
                    PM_ACC(14);
                }
                const int w = min(a.eb, ncols_valid - nb);
                if (o >= 0 && w > 0)
                    umma::bulk_reduce_add_f32(a.out + (size_t)o * a.Cout + col0 + nb, T, (uint32_t)w * 4);
                umma::bulk_commit();
            }
            // next item's row data: issued after this item's proxy fences (a fence waits for the
            // thread's outstanding loads), consumed after the next accumulator wait
            if (n + 1 < n_items) meta(it, o_next, imp_next);
        }
        umma::bulk_wait_all();
        if (a.dbg && blockIdx.x == 0 && tid == 0) {
            g_pm_dbg[6] += (unsigned long long)(clock64() - te0);
            PM_FLUSH(7);
            PM_FLUSH(8);
            PM_FLUSH(13);
            PM_FLUSH(14);
        }
    } else if (warp < kLoadWarp0) {
        // ================================================================ converters
        // raw tile (fp32 bits; the tensor core reads them as tf32 = hi, ignoring the 13 low
        // mantissa bits) -> lo = x - hi tile.  No global loads here, so the proxy fence is cheap.
        const int ct = tid - kEpiWarps * 32;  // 0..127
        const int kq = ct & 3, rsub = ct >> 2;  // rows rsub + 32 j, j = 0..3
        const uint32_t off0 = (uint32_t)(rsub >> 3) * kA_SBO + (uint32_t)kq * kA_LBO + (uint32_t)(rsub & 7) * 16;
        int r = 0, rph = 0, q = 0, qph = 0;
        const long long tc0 = a.dbg ? clock64() : 0;
        for (int step = 0; step < total_steps; ++step) {
            {
                PM_T0();
                mbar_wait(&raw_full[r], rph);
                PM_ACC(1);
            }
            const uint8_t* src = sRaw + (size_t)r * kATileBytes + off0;
            float4 l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 x = *reinterpret_cast<const float4*>(src + (uint32_t)j * 4 * kA_SBO);
                l[j].x = x.x - umma::tf32_hi(x.x); l[j].y = x.y - umma::tf32_hi(x.y);
                l[j].z = x.z - umma::tf32_hi(x.z); l[j].w = x.w - umma::tf32_hi(x.w);
            }
            if (step >= Q) {
                PM_T0();
                mbar_wait(&lo_empty[q], qph ^ 1);
                PM_ACC(12);
            }
            uint8_t* dst = sLo + (size_t)q * kATileBytes + off0;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(dst + (uint32_t)j * 4 * kA_SBO) = l[j];
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
        if (a.dbg && blockIdx.x == 0 && ct == 0) {
            g_pm_dbg[0] += (unsigned long long)(clock64() - tc0);
            PM_FLUSH(1);
            PM_FLUSH(12);
        }
    } else if (warp < kMmaWarp) {
        // ================================================================ loaders
        // 16-byte cp.async pieces straight into the K-major raw tile; a thread never waits for
        // its data, so up to R - 2 stages (8 KB each) are in flight per SM.
        const int lw = warp - kLoadWarp0;
        const int kq = lane & 3, rl = lane >> 2;
        // rows lw * 64 + 8 j + rl, j = 0..7
        const uint32_t off_base = (uint32_t)(lw * 8) * kA_SBO + (uint32_t)kq * kA_LBO + (uint32_t)rl * 16;
        auto pins_of = [&](const ItemIter& qi, int (&pin)[8]) {
            const int cnt = (int)min((long long)TM, qi.len - (long long)qi.t * TM);
            const long long s0 = qi.base + (long long)qi.t * TM;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int row = lw * 64 + 8 * j + rl;
                pin[j] = row < cnt ? __ldg(a.p_in + s0 + row) : -1;
            }
        };
        int pin[8], pin_next[8];
        ItemIter nx = it;
        if (n_items > 0) pins_of(it, pin);
        if (n_items > 1) {
            iter_next(a, nx);
            pins_of(nx, pin_next);
        }
        int r = 0, rph = 0, step = 0;
        const long long tl0 = a.dbg ? clock64() : 0;
        for (int n = 0; n < n_items; ++n) {
            const int k0 = (it.p / a.nparts) * a.kc + kq * 4;
            for (int c = 0; c < nch; ++c, ++step) {
                if (step >= R) {
                    PM_T0();
                    mbar_wait(&raw_empty[r], rph ^ 1);
                    PM_ACC(11);
                }
                const uint32_t dst = umma::smem_u32(sRaw + (size_t)r * kATileBytes) + off_base;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool ok = pin[j] >= 0;
                    const float* src = ok ? a.x + (size_t)pin[j] * a.Cin + k0 + c * KC : a.x;
                    umma::cp_async16_cg(dst + (uint32_t)j * kA_SBO, src, ok ? 16u : 0u);
                }
                umma::cp_async_arrive_noinc(&raw_full[r]);
                if (++r == R) {
                    r = 0;
                    rph ^= 1;
                }
            }
            iter_next(a, it);
            if (n + 1 < n_items) {
#pragma unroll
                for (int j = 0; j < 8; ++j) pin[j] = pin_next[j];
                if (n + 2 < n_items) {
                    iter_next(a, nx);
                    pins_of(nx, pin_next);
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (a.dbg && blockIdx.x == 0 && tid == kLoadWarp0 * 32) {
            g_pm_dbg[10] += (unsigned long long)(clock64() - tl0);
            PM_FLUSH(11);
        }
    } else if (lane == 0) {
        // ================================================================ MMA issuer + filter loader
        const long long tm0 = a.dbg ? clock64() : 0;
        const uint32_t idesc = umma::make_idesc_tf32(128, nc), idesc2 = umma::make_idesc_tf32(128, 2 * nc);
        const int chunks_total = (a.Cin + KC - 1) / KC;
        const uint64_t d_raw0 = umma::desc_base(kA_LBO, kA_SBO) + (umma::smem_u32(sRaw) >> 4);
        const uint64_t d_lo0 = umma::desc_base(kA_LBO, kA_SBO) + (umma::smem_u32(sLo) >> 4);
        const uint64_t d_b0 = umma::desc_base(kB_LBO, kB_SBO) + (umma::smem_u32(sB) >> 4);
        int cur_slot = -1, cur_p = -1;
        uint32_t b_phase = 0;
        int r = 0, q = 0, qph = 0;
        for (int n = 0; n < n_items; ++n) {
            const int buf = n & 1, use = n >> 1;
            const int slot = it.g < a.num_blocks ? 0 : 1 + (it.g - a.num_blocks) % (a.K - 1), p = it.p;
            bool wait_b = false;
            if (slot != cur_slot || p != cur_p) {
                // every MMA issued so far has read the old filter part: wait for the last commit
                if (n > 0) mbar_wait(&acc_full[(n - 1) & 1], ((n - 1) >> 1) & 1);
                const int kp = p / a.nparts, np = p % a.nparts;
                const float* src = a.wp + ((size_t)slot * chunks_total + (size_t)kp * nch) * 2 * a.n_pad * KC +
                                   (size_t)np * nc * KC;
                umma::mbar_arrive_expect_tx(&b_full, (uint32_t)nch * b_chunk_bytes);
                for (int c = 0; c < nch; ++c) {
                    uint8_t* dst = sB + (size_t)c * b_chunk_bytes;
                    const float* s_hi = src + (size_t)c * 2 * a.n_pad * KC;
                    bulk_copy_g2s(dst, s_hi, b_chunk_bytes / 2, &b_full);
                    bulk_copy_g2s(dst + b_chunk_bytes / 2, s_hi + (size_t)a.n_pad * KC, b_chunk_bytes / 2, &b_full);
                }
                cur_slot = slot;
                cur_p = p;
                wait_b = true;
            }
            iter_next(a, it);
            if (use > 0) {
                PM_T0();
                mbar_wait(&acc_empty[buf], (use - 1) & 1);
                PM_ACC(4);
            }
            const uint32_t t_main = tmem + (uint32_t)buf * 2 * nc, t_corr = t_main + nc;
            for (int c = 0; c < nch; ++c) {
                {
                    PM_T0();
                    mbar_wait(&lo_full[q], qph);  // implies the raw tile has landed (the converters read it)
                    PM_ACC(3);
                }
                if (wait_b) {
                    PM_T0();
                    mbar_wait(&b_full, b_phase);
                    PM_ACC(5);
                    b_phase ^= 1;
                    wait_b = false;
                }
                // the raw tile was written through the generic proxy (cp.async); the converters'
                // fence.proxy.async lies on the causality path loader -> converter -> this thread
                umma::tc_fence_after();
                PM_T0();
                const uint64_t a_hi0 = d_raw0 + (uint64_t)(r * (int)(kATileBytes >> 4));
                const uint64_t a_lo0 = d_lo0 + (uint64_t)(q * (int)(kATileBytes >> 4));
                const uint64_t b_hi0 = d_b0 + (uint64_t)(c * (int)(b_chunk_bytes >> 4));
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t oa = (uint64_t)(ks * 2 * (kA_LBO >> 4)), ob = (uint64_t)(ks * 2 * (kB_LBO >> 4));
                    // A_hi x [B_hi | B_lo] -> (main | correction) in one MMA of width 2 nc (sparse_conv_tc.cu)
                    umma::mma_tf32(t_main, a_hi0 + oa, b_hi0 + ob, idesc2, c > 0 || ks > 0);
                    umma::mma_tf32_acc(t_corr, a_lo0 + oa, b_hi0 + ob, idesc);
                }
                umma::mma_commit(&raw_empty[r]);
                umma::mma_commit(&lo_empty[q]);
                PM_ACC(15);
                if (++r == R) r = 0;
                if (++q == Q) {
                    q = 0;
                    qph ^= 1;
                }
            }
            umma::mma_commit(&acc_full[buf]);
        }
        if (a.dbg && blockIdx.x == 0) {
            g_pm_dbg[2] += (unsigned long long)(clock64() - tm0);
            g_pm_dbg[9] += n_items;
            PM_FLUSH(3);
            PM_FLUSH(4);
            PM_FLUSH(5);
            PM_FLUSH(15);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) umma::tmem_dealloc(tmem, ncols);
}

static bool g_pm_enabled = false;  // see DESIGN.md: bounded by the L2 reduction rate like the per-tile kernel
static int g_pm_debug = 0;
void sparse_conv_pm_enable(bool on) { g_pm_enabled = on; }
void sparse_conv_pm_debug(int on) { g_pm_debug = on; }

bool sparse_conv_pm_supported(const ConvPlan& P, int Cin, int Cout) {
    return g_pm_enabled && P.g_tile0.size() > 0 && Cin % KC == 0 && (Cin <= 128 || Cin % 128 == 0) && Cout % 32 == 0 &&
           Cout <= 256;
}

void sparse_conv_pm_tiles(const ConvPlan& P, const float* x, const float* wp, int Cin, int Cout, const float* imp_in,
                          const float* imp_entry, int imp_col, float* out, cudaStream_t s) {
    PmArgs a;
    a.x = x;
    a.wp = wp;
    a.p_in = P.p_in.get();
    a.p_out = P.p_out.get();
    a.perm = P.perm.get();
    a.g_begin = P.g_begin.get();
    a.g_tile0 = P.g_tile0.get();
    a.imp_in = imp_in;
    a.imp_entry = imp_entry;
    a.out = out;
    a.G = P.G;
    a.K = P.K;
    a.num_blocks = P.num_blocks;
    a.Cin = Cin;
    a.Cout = Cout;
    a.n_pad = Cout;
    a.imp_col = (imp_in || imp_entry) ? imp_col : Cout;
    a.kc = std::min(Cin, 128);
    a.nc = std::min(a.n_pad, 128);
    a.nparts = a.n_pad / a.nc;
    a.parts = (Cin / a.kc) * a.nparts;
    a.nch = a.kc / KC;
    // shared memory: resident filter part | raw ring | lo ring | epilogue staging rows
    const size_t budget = 226 * 1024;
    const size_t b_bytes = (size_t)a.nch * 2 * a.nc * KC * 4;
    a.Q = b_bytes > 64 * 1024 ? 2 : kMaxLo;
    a.eb = std::min(a.nc / 2, 32);  // columns one epilogue thread stages per bulk reduction
    const size_t t_bytes = (size_t)kEpiWarps * 32 * (a.eb + 4) * sizeof(float);
    {
        const size_t fixed = b_bytes + (size_t)a.Q * kATileBytes + t_bytes;
        a.R = fixed >= budget ? 0 : (int)std::min<size_t>(kMaxRaw, (budget - fixed) / kATileBytes);
    }
    ASRB_REQUIRE(a.R >= 3, "sparse_conv_pm: shared-memory budget");
    a.off_raw = (uint32_t)b_bytes;
    a.off_lo = (uint32_t)(b_bytes + (size_t)a.R * kATileBytes);
    a.off_t = (uint32_t)(a.off_lo + (size_t)a.Q * kATileBytes);
    const size_t smem = a.off_t + t_bytes;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        ASRB_CUDA(cudaGetDevice(&dev));
        ASRB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    ASRB_CUDA(cudaFuncSetAttribute(sparse_conv_pm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    char label[96];
    snprintf(label, sizeof(label), "sparse_conv_tile/pm K%d %dx%d E%lld", P.K, Cin, Cout, (long long)P.E);
    ProfileScope prof(label, s, 2.0 * (double)P.E * Cin * Cout);
    const long long max_items = (long long)P.max_tiles * a.parts;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(n_sm, max_items));
    a.dbg = g_pm_debug;
    if (g_pm_debug == 1) {
        unsigned long long z[16] = {0};
        ASRB_CUDA(cudaMemcpyToSymbol(g_pm_dbg, z, sizeof(z)));
    }
    sparse_conv_pm_kernel<<<grid, kThreads, smem, s>>>(a);
    ASRB_CHECK_LAUNCH();
    if (g_pm_debug == 1) {
        unsigned long long h[16];
        ASRB_CUDA(cudaStreamSynchronize(s));
        ASRB_CUDA(cudaMemcpyFromSymbol(h, g_pm_dbg, sizeof(h)));
        fprintf(stderr, "[pm] %s R %d Q %d eb %d items %llu | load total %llu wait_raw_empty %llu | cvt total %llu "
                        "wait_raw_full %llu wait_lo_empty %llu | mma total %llu wait_lo_full %llu wait_acc_empty %llu "
                        "wait_b %llu issue %llu | epi total %llu wait_acc_full %llu wait_stage %llu tmem_ld %llu fence %llu (cycles, CTA 0)\n",
                label, a.R, a.Q, a.eb, h[9], h[10], h[11], h[0], h[1], h[12], h[2], h[3], h[4], h[5], h[15], h[6], h[7], h[8], h[13], h[14]);
    }
}

}  // namespace asrb
