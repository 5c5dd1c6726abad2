// Multi-radius neighbour search: for every query q_i, all points p with
// |p - q_i|^2 < r_i^2, ascending by squared distance.
//
// Replaces the Open3D NearestNeighborSearch::MultiRadiusIndex/MultiRadiusSearch
// calls (nanoflann KD-tree + TBB) made by
// ComputeAggregationNeighborsAndScaleCompatibility (reference
// cpp/lib/nsearch.cpp:107-162; Python twin models/v0/datareader.py:776-795).
//
// GPU design: points are binned on a 2^21-per-axis grid over their bounding cube
// and sorted by Morton code once (radix sort), so every cubic cell of any
// power-of-two size is one contiguous window of the sorted array.  One warp owns
// a query: 16 lanes binary-search the window bounds of the <= 2x2x2 cells (edge
// >= 2r) that cover the query ball, then all 32 lanes stream the concatenated
// windows with coalesced 16-byte loads, test the strict squared distance and
// compact hits with ballot/popc.  count -> scan -> fill with the per-row rank sort by
// (d2, index) fused in (shared memory for rows of <= 32 hits).
//
// Squared distances are evaluated as ((dx*dx) + dy*dy) + dz*dz in fp32 without
// FMA contraction, the order nanoflann's L2 adaptor uses for dim 3, so that the
// strict `<` test resolves boundary points identically.
#include "hash.cuh"
#include "internal.h"
#include "prims.cuh"
#include "profile.cuh"
#include "search.h"

namespace asrb {

constexpr int kGridBits = 21;
constexpr float kGridMax = 2097151.0f;  // 2^21 - 1

struct BinFrame {
    float origin[3];
    float inv_h;
};

__device__ __forceinline__ int bin_coord(float x, float origin, float inv_h) {
    float v = floorf(__fmul_rn(__fsub_rn(x, origin), inv_h));
    v = fminf(fmaxf(v, 0.0f), kGridMax);  // monotone in x; NaN -> 0
    return (int)v;
}

__device__ __forceinline__ unsigned ordered_bits(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float from_ordered_bits(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// bounding box of the points: block reduction + 6 atomics per block
__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ pts, long long n, unsigned* __restrict__ mn, unsigned* __restrict__ mx) {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        for (int a = 0; a < 3; ++a) {
            const float v = pts[3 * i + a];
            if (v == v) {  // ignore NaN
                const unsigned b = ordered_bits(v);
                lo[a] = min(lo[a], b);
                hi[a] = max(hi[a], b);
            }
        }
    }
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 3; ++a) {
            atomicMin(mn + a, lo[a]);
            atomicMax(mx + a, hi[a]);
        }
}

__global__ void __launch_bounds__(256)
point_code_kernel(const float* __restrict__ pts, long long n, BinFrame f, Key* __restrict__ codes,
                  uint32_t* __restrict__ order) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = bin_coord(pts[3 * i], f.origin[0], f.inv_h);
    const int y = bin_coord(pts[3 * i + 1], f.origin[1], f.inv_h);
    const int z = bin_coord(pts[3 * i + 2], f.origin[2], f.inv_h);
    codes[i] = morton3(x, y, z);
    order[i] = (uint32_t)i;
}

// sorted position -> (x, y, z, original index) in one 16-byte record
__global__ void __launch_bounds__(256)
gather_points_kernel(const float* __restrict__ pts, const uint32_t* __restrict__ order, long long n,
                     float4* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = order[i];
    out[i] = make_float4(pts[3 * (size_t)j], pts[3 * (size_t)j + 1], pts[3 * (size_t)j + 2], __uint_as_float(j));
}

constexpr int kWarpsPerBlock = 8;

// ------------------------------------------------------------------ cell index
// For every power-of-two cell size that some query needs, the non-empty cells of
// the Morton-sorted point array are entered in a hash table
//     (cell code | level marker)  ->  [begin, end) window of the sorted points,
// so a query resolves each covering cell with ~1.5 probes.  Position j of the
// sorted array starts a new cell at shift sh iff the highest bit in which its
// code differs from its predecessor's lies at or above 3*sh, so one pass over
// the points serves all levels.
__device__ __forceinline__ Key cell_key_at(Key code, int sh) {
    return (code >> (3 * sh)) | (Key(1) << (3 * (kGridBits - sh)));
}
__device__ __forceinline__ int boundary_top_shift(const Key* __restrict__ codes, long long j) {
    if (j == 0) return kGridBits;
    const Key d = codes[j] ^ codes[j - 1];
    return d ? (63 - __clzll((long long)d)) / 3 : -1;
}

// ball -> conservative integer box [lo, hi] on the 2^21 grid and the smallest
// shift at which it spans <= 3 cells per axis
__device__ __forceinline__ int query_box(const float* __restrict__ q, float r, const BinFrame& f, int lo[3], int hi[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float pad = __fmaf_rn(r, 1.00001f, fabsf(q[a]) * 4e-7f);  // a few ulps of the coordinates
        lo[a] = bin_coord(q[a] - pad, f.origin[a], f.inv_h);
        hi[a] = bin_coord(q[a] + pad, f.origin[a], f.inv_h);
    }
    const int span = max(max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]) + 1;
    int sh = span <= 3 ? 0 : max(0, 30 - __clz(span));  // 2^(sh+2) > span: a first guess, then refine
    while (sh < kGridBits && (((hi[0] >> sh) - (lo[0] >> sh)) > 2 || ((hi[1] >> sh) - (lo[1] >> sh)) > 2 ||
                              ((hi[2] >> sh) - (lo[2] >> sh)) > 2))
        ++sh;
    return sh;
}

__global__ void __launch_bounds__(256)
query_levels_kernel(const float* __restrict__ queries, const float* __restrict__ radii, long long nq, BinFrame f,
                    unsigned* __restrict__ level_mask) {
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    unsigned m = 0;
    if (q < nq) {
        const float r = radii[q];
        if (r >= 0.0f) {
            int lo[3], hi[3];
            m = 1u << query_box(queries + 3 * q, r, f, lo, hi);
        }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(level_mask, m);
}

__global__ void __launch_bounds__(256)
cell_count_kernel(const Key* __restrict__ codes, long long n, unsigned levels, unsigned long long* __restrict__ total) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    unsigned c = 0;
    if (j < n) {
        const int top = boundary_top_shift(codes, j);
        if (top >= 0) c = __popc(levels & ((2u << top) - 1u));
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, (unsigned long long)c);
}

__global__ void __launch_bounds__(256)
cell_insert_kernel(const Key* __restrict__ codes, long long n, unsigned levels, HashEntry* __restrict__ e, uint32_t mask) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int top = boundary_top_shift(codes, j);
    if (top < 0) return;
    unsigned lv = levels & ((2u << top) - 1u);
    const Key code = codes[j];
    while (lv) {
        const int sh = __ffs(lv) - 1;
        lv &= lv - 1;
        const Key k = cell_key_at(code, sh);
        uint32_t s = hash_key(k) & mask;
        for (;;) {
            const Key prev = atomicCAS(&e[s].key, kNoKey, k);
            if (prev == kNoKey) {
                e[s].val = j;  // begin; the end is filled in by cell_end_kernel
                break;
            }
            s = (s + 1) & mask;
        }
    }
}

// the cell that ends at position j (exclusive) is the one containing point j-1
__global__ void __launch_bounds__(256)
cell_end_kernel(const Key* __restrict__ codes, long long n, unsigned levels, HashEntry* __restrict__ e, uint32_t mask) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x + 1;
    if (j > n) return;
    const int top = j == n ? kGridBits : boundary_top_shift(codes, j);
    if (top < 0) return;
    unsigned lv = levels & ((2u << top) - 1u);
    const Key code = codes[j - 1];
    while (lv) {
        const int sh = __ffs(lv) - 1;
        lv &= lv - 1;
        const Key k = cell_key_at(code, sh);
        uint32_t s = hash_key(k) & mask;
        while (e[s].key != k) s = (s + 1) & mask;
        e[s].val |= (long long)j << 32;
    }
}

constexpr int kStash = 64;

// rows of <= kStash hits: rank sort of the stashed (d2, index) keys straight into the result
// (a lane holds keys `lane` and `lane + 32`)
__global__ void __launch_bounds__(256)
stash_rows_kernel(const unsigned long long* __restrict__ stash, const int64_t* __restrict__ splits, long long nq,
                  int32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
    const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const int64_t b = splits[q];
    const int n = (int)(splits[q + 1] - b);
    if (n == 0 || n > kStash) return;
    const unsigned long long m0 = lane < n ? stash[q * kStash + lane] : ~0ULL;
    const unsigned long long m1 = lane + 32 < n ? stash[q * kStash + 32 + lane] : ~0ULL;
    int r0 = 0, r1 = 0;
    for (int j = 0; j < min(n, 32); ++j) {
        const unsigned long long k = __shfl_sync(0xffffffffu, m0, j);
        r0 += k < m0 ? 1 : 0;
        r1 += k < m1 ? 1 : 0;
    }
    for (int j = 32; j < n; ++j) {
        const unsigned long long k = __shfl_sync(0xffffffffu, m1, j - 32);
        r0 += k < m0 ? 1 : 0;
        r1 += k < m1 ? 1 : 0;
    }
    if (lane < n) {
        out_idx[b + r0] = (int32_t)(unsigned)m0;
        out_d2[b + r0] = __uint_as_float((unsigned)(m0 >> 32));
    }
    if (lane + 32 < n) {
        out_idx[b + r1] = (int32_t)(unsigned)m1;
        out_d2[b + r1] = __uint_as_float((unsigned)(m1 >> 32));
    }
}

template <bool FILL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ball_query_kernel(const KeyTableView cells, const float4* __restrict__ spts, BinFrame f,
                  const float* __restrict__ queries, const float* __restrict__ radii, long long nq,
                  int32_t* __restrict__ counts, const int64_t* __restrict__ splits,
                  unsigned long long* __restrict__ out_keys, int32_t* __restrict__ out_idx,
                  float* __restrict__ out_d2, unsigned long long* __restrict__ stash) {
    // count pass (FILL = false): the first kStash hits of every query are also kept in `stash`
    // ([nq][kStash] keys), so rows of <= kStash hits (nearly all) are finished by
    // stash_rows_kernel without streaming the candidates again; the fill pass (FILL = true)
    // then only redoes the longer rows (skip_small: stash != nullptr).
    __shared__ unsigned s_begin[kWarpsPerBlock][32];
    __shared__ int s_pre[kWarpsPerBlock][33];
    __shared__ unsigned long long s_hit[FILL ? kWarpsPerBlock : 1][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q = blockIdx.x * (long long)kWarpsPerBlock + warp;
    if (q >= nq) return;
    const float qx = queries[3 * q], qy = queries[3 * q + 1], qz = queries[3 * q + 2];
    if (FILL && stash && splits[q + 1] - splits[q] <= kStash) return;
    const float r = radii[q];
    const float r2 = __fmul_rn(r, r);
    int len = 0;
    unsigned begin = 0;
    if (r >= 0.0f) {
        int lo[3], hi[3];
        const float qq[3] = {qx, qy, qz};
        const int sh = query_box(qq, r, f, lo, hi);
        if (lane < 27) {
            const int cx = (lo[0] >> sh) + lane % 3, cy = (lo[1] >> sh) + (lane / 3) % 3, cz = (lo[2] >> sh) + lane / 9;
            if (cx <= (hi[0] >> sh) && cy <= (hi[1] >> sh) && cz <= (hi[2] >> sh)) {
                const Key k = morton3(cx, cy, cz) | (Key(1) << (3 * (kGridBits - sh)));
                const long long v = table_find(cells, k);
                if (v >= 0) {
                    begin = (unsigned)v;
                    len = (int)((unsigned long long)v >> 32) - (int)begin;
                }
            }
        }
    }
    int pre = len;  // inclusive prefix over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, d);
        if (lane >= d) pre += v;
    }
    s_begin[warp][lane] = begin;
    s_pre[warp][lane + 1] = pre;
    if (lane == 0) s_pre[warp][0] = 0;
    __syncwarp();
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    int found = 0;
    const int64_t out_base = FILL ? splits[q] : 0;
    // FILL: rows of <= 32 hits (the rule) are collected in shared memory and rank-sorted by
    // (d2, index) right here; longer rows go through `out_keys` and are sorted below
    const int row_n = FILL ? (int)(splits[q + 1] - out_base) : 0;
    const bool small = row_n <= 32;
    for (int t0 = 0; t0 < total; t0 += 32) {
        const int t = t0 + lane;
        bool hit = false;
        unsigned long long key = 0;
        if (t < total) {
            int c = 0;  // last cell with s_pre[c] <= t
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (c + step < 32 && s_pre[warp][c + step] <= t) c += step;
            const float4 p = __ldg(spts + s_begin[warp][c] + (unsigned)(t - s_pre[warp][c]));
            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            hit = d2 < r2;
            key = ((unsigned long long)__float_as_uint(d2) << 32) | __float_as_uint(p.w);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) {
            const int pos = found + __popc(m & ((1u << lane) - 1));
            if (small) s_hit[warp][pos] = key;
            else out_keys[out_base + pos] = key;
        }
        if (!FILL && stash && hit) {
            const int pos = found + __popc(m & ((1u << lane) - 1));
            if (pos < kStash) stash[q * kStash + pos] = key;
        }
        found += __popc(m);
    }
    if (!FILL) {
        if (lane == 0) counts[q] = found;
        return;
    }
    __syncwarp();
    if (small) {
        const unsigned long long mine = lane < row_n ? s_hit[warp][lane] : ~0ULL;
        int rank = 0;
        for (int j = 0; j < row_n; ++j) rank += __shfl_sync(0xffffffffu, mine, j) < mine ? 1 : 0;
        if (lane < row_n) {
            out_idx[out_base + rank] = (int32_t)(unsigned)mine;
            out_d2[out_base + rank] = __uint_as_float((unsigned)(mine >> 32));
        }
    } else {
        __threadfence_block();  // this warp's own key writes, read back by all its lanes
        for (int i = lane; i < row_n; i += 32) {
            const unsigned long long mine = out_keys[out_base + i];
            int rank = 0;
            for (int j = 0; j < row_n; ++j) rank += out_keys[out_base + j] < mine ? 1 : 0;
            out_idx[out_base + rank] = (int32_t)(unsigned)mine;
            out_d2[out_base + rank] = __uint_as_float((unsigned)(mine >> 32));
        }
    }
}

// ---- rows of more than kStash hits (1.2 % of the rows, 8 % of the pairs on the bench cloud, up to ~1100 hits) ----
// A warp per such row (the first version) streamed the row's candidates with 32 lanes and rank-sorted n keys in
// n^2 / 32 steps through global memory: 3.2 ms for 42 k rows.  Now: the rows are listed, a BLOCK takes a row:
// 256 lanes stream the windows, hits are appended to shared memory (any order: the keys (d2, index) are
// distinct, the rank sort fixes the result), and the rank of each key is counted against the shared copy.
constexpr int kHeavyKeys = 2048;  // keys of a row kept in shared memory; beyond that they go through `gkeys`

__global__ void __launch_bounds__(256)
heavy_list_kernel(const int64_t* __restrict__ splits, long long nq, int32_t* __restrict__ list, int* __restrict__ count) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (splits[q + 1] - splits[q] > kStash) list[atomicAdd(count, 1)] = (int32_t)q;
}

__global__ void __launch_bounds__(256)
ball_query_heavy_kernel(const KeyTableView cells, const float4* __restrict__ spts, BinFrame f,
                        const float* __restrict__ queries, const float* __restrict__ radii,
                        const int64_t* __restrict__ splits, const int32_t* __restrict__ list,
                        const int* __restrict__ count, unsigned long long* __restrict__ gkeys,
                        int32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
    __shared__ unsigned long long s_key[kHeavyKeys];
    __shared__ unsigned s_begin[32];
    __shared__ int s_pre[33];
    __shared__ int s_n;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nheavy = *count;
    for (int h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const long long q = list[h];
        const float qx = queries[3 * q], qy = queries[3 * q + 1], qz = queries[3 * q + 2];
        const float r = radii[q];
        const float r2 = __fmul_rn(r, r);
        const int64_t out_base = splits[q];
        const int row_n = (int)(splits[q + 1] - out_base);
        __syncthreads();  // the previous row's shared data is no longer read
        if (tid < 32) {
            int len = 0;
            unsigned begin = 0;
            int lo[3], hi[3];
            const float qq[3] = {qx, qy, qz};
            const int sh = query_box(qq, r, f, lo, hi);
            if (lane < 27) {
                const int cx = (lo[0] >> sh) + lane % 3, cy = (lo[1] >> sh) + (lane / 3) % 3, cz = (lo[2] >> sh) + lane / 9;
                if (cx <= (hi[0] >> sh) && cy <= (hi[1] >> sh) && cz <= (hi[2] >> sh)) {
                    const Key k = morton3(cx, cy, cz) | (Key(1) << (3 * (kGridBits - sh)));
                    const long long v = table_find(cells, k);
                    if (v >= 0) {
                        begin = (unsigned)v;
                        len = (int)((unsigned long long)v >> 32) - (int)begin;
                    }
                }
            }
            int pre = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, d);
                if (lane >= d) pre += v;
            }
            s_begin[lane] = begin;
            s_pre[lane + 1] = pre;
            if (lane == 0) {
                s_pre[0] = 0;
                s_n = 0;
            }
        }
        __syncthreads();
        const int total = s_pre[32];
        for (int t = tid; t < total; t += 256) {
            int c = 0;  // last cell with s_pre[c] <= t
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (c + step < 32 && s_pre[c + step] <= t) c += step;
            const float4 p = __ldg(spts + s_begin[c] + (unsigned)(t - s_pre[c]));
            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 < r2) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | __float_as_uint(p.w);
                const int pos = atomicAdd(&s_n, 1);
                if (pos < kHeavyKeys) s_key[pos] = key;
                else gkeys[out_base + pos] = key;
            }
        }
        __threadfence_block();
        __syncthreads();
        const int ns = min(row_n, kHeavyKeys);
        for (int i = tid; i < row_n; i += 256) {
            const unsigned long long mine = i < kHeavyKeys ? s_key[i] : gkeys[out_base + i];
            int rank = 0;
            for (int j = 0; j < ns; ++j) rank += s_key[j] < mine ? 1 : 0;
            for (int j = kHeavyKeys; j < row_n; ++j) rank += gkeys[out_base + j] < mine ? 1 : 0;
            out_idx[out_base + rank] = (int32_t)(unsigned)mine;
            out_d2[out_base + rank] = __uint_as_float((unsigned)(mine >> 32));
        }
    }
}

// compat = (min(s_v, 2 r_p) / max(s_v, 2 r_p))^2   (nsearch.cpp:149-161)
__global__ void __launch_bounds__(256)
scale_compat_kernel(const float* __restrict__ sizes, const float* __restrict__ radii, const int32_t* __restrict__ idx,
                    const int64_t* __restrict__ splits, long long nq, float* __restrict__ out) {
    const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float a = sizes[q];
    const int64_t e = splits[q + 1];
    for (int64_t j = splits[q] + lane; j < e; j += 32) {
        const float b = 2.0f * __ldg(radii + idx[j]);
        const float t = fminf(a, b) / fmaxf(a, b);
        out[j] = __fmul_rn(t, t);
    }
}

void search_prepare(Search& S, const float* d_points, int64_t n, const float* d_queries, const float* d_radii,
                    int64_t nq, const float* h_frame, cudaStream_t s) {
    S.n = n;
    S.nq = nq;
    S.queries = d_queries;
    S.radii = d_radii;
    if (h_frame) {
        // caller-supplied binning frame (origin, finest cell size): the pipeline passes
        // the octree's, which makes the cells coincide with the voxels being queried
        for (int a = 0; a < 3; ++a) S.frame_origin[a] = h_frame[a];
        S.frame_inv_h = 1.0f / h_frame[3];
    } else {
        // bounding cube of the points
        DevBuf<unsigned> mm(6, s);
        ASRB_CUDA(cudaMemsetAsync(mm.get(), 0xff, 3 * sizeof(unsigned), s));
        ASRB_CUDA(cudaMemsetAsync(mm.get() + 3, 0, 3 * sizeof(unsigned), s));
        unsigned h[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0};
        if (n > 0) {
            const unsigned blocks = (unsigned)std::min<size_t>(grid_for(n, 256), 148 * 8);
            bbox_kernel<<<blocks, 256, 0, s>>>(d_points, n, mm.get(), mm.get() + 3);
            ASRB_CHECK_LAUNCH();
            ASRB_CUDA(cudaMemcpyAsync(h, mm.get(), sizeof(h), cudaMemcpyDeviceToHost, s));
            ASRB_CUDA(cudaStreamSynchronize(s));
        }
        float edge = 0.f;
        for (int a = 0; a < 3; ++a) {
            const float lo = h[a] == 0xffffffffu ? 0.f : from_ordered_bits(h[a]);
            const float hi = h[3 + a] == 0 ? 0.f : from_ordered_bits(h[3 + a]);
            S.frame_origin[a] = lo;
            edge = std::max(edge, hi - lo);
        }
        if (!(edge > 0.f)) edge = 1.f;
        S.frame_inv_h = (float)(2097152.0 / ((double)edge * 1.000001));
    }
    BinFrame f{{S.frame_origin[0], S.frame_origin[1], S.frame_origin[2]}, S.frame_inv_h};

    S.codes.alloc((size_t)n, s);
    S.spts.alloc((size_t)n, s);
    DevBuf<uint32_t> order((size_t)n, s);
    if (n > 0) {
        ProfileScope prof("search_point_sort", s);
        point_code_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_points, n, f, S.codes.get(), order.get());
        ASRB_CHECK_LAUNCH();
        sort_pairs_u64_u32(S.codes.get(), order.get(), (size_t)n, s, 63);
        gather_points_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_points, order.get(), n, (float4*)S.spts.get());
        ASRB_CHECK_LAUNCH();
    }
    // cell index for the cell sizes the queries need
    {
        ProfileScope prof("search_cell_index", s);
        DevBuf<unsigned long long> scal(2, s);
        ASRB_CUDA(cudaMemsetAsync(scal.get(), 0, 2 * sizeof(unsigned long long), s));
        unsigned* d_levels = reinterpret_cast<unsigned*>(scal.get() + 1);
        if (nq > 0) {
            query_levels_kernel<<<grid_for(nq, 256), 256, 0, s>>>(d_queries, d_radii, nq, f, d_levels);
            ASRB_CHECK_LAUNCH();
        }
        unsigned levels = d2h_scalar(d_levels, s);
        levels &= (2u << kGridBits) - 1u;
        if (n > 0 && levels) {
            cell_count_kernel<<<grid_for(n, 256), 256, 0, s>>>(S.codes.get(), n, levels, scal.get());
            ASRB_CHECK_LAUNCH();
        }
        const unsigned long long ncells = d2h_scalar(scal.get(), s);
        size_t cap = 64;
        while (cap < 2 * ncells) cap <<= 1;
        S.cells.mask = (uint32_t)(cap - 1);
        S.cells.entries.alloc(cap, s);
        ASRB_CUDA(cudaMemsetAsync(S.cells.entries.get(), 0xff, cap * sizeof(HashEntry), s));  // key = kNoKey, val = -1
        if (n > 0 && levels) {
            cell_insert_kernel<<<grid_for(n, 256), 256, 0, s>>>(S.codes.get(), n, levels, S.cells.entries.get(),
                                                                S.cells.mask);
            ASRB_CHECK_LAUNCH();
            cell_end_kernel<<<grid_for(n, 256), 256, 0, s>>>(S.codes.get(), n, levels, S.cells.entries.get(),
                                                             S.cells.mask);
            ASRB_CHECK_LAUNCH();
        }
    }
    // count pass (stashing the first kStash hits per query when the results will be asked for)
    DevBuf<int32_t> counts((size_t)nq, s);
    if (S.want_fill) S.stash.alloc((size_t)nq * kStash, s);
    S.splits.alloc((size_t)nq + 1, s);
    if (nq > 0) {
        ProfileScope prof("ball_query_count", s);
        ball_query_kernel<false><<<grid_for(nq, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                S.cells.view(), (const float4*)S.spts.get(), f, d_queries, d_radii, nq, counts.get(), nullptr, nullptr, nullptr,
                nullptr, S.stash.get());
        ASRB_CHECK_LAUNCH();
    }
    exclusive_sum_i32_to_i64(counts.get(), S.splits.get(), (size_t)nq, s);
    S.num_pairs = d2h_scalar(S.splits.get() + nq, s);
}

void search_fill(Search& S, int32_t* d_idx, float* d_d2, int64_t* d_splits, cudaStream_t s) {
    BinFrame f{{S.frame_origin[0], S.frame_origin[1], S.frame_origin[2]}, S.frame_inv_h};
    ASRB_CUDA(cudaMemcpyAsync(d_splits, S.splits.get(), ((size_t)S.nq + 1) * sizeof(int64_t),
                              cudaMemcpyDeviceToDevice, s));
    if (S.nq == 0 || S.num_pairs == 0) return;
    DevBuf<unsigned long long> keys((size_t)S.num_pairs, s);
    ProfileScope prof("ball_query_fill_sort", s);
    if (S.stash.size()) {
        // rows of <= kStash hits straight from the stash, the longer ones by a block each
        stash_rows_kernel<<<grid_for((size_t)S.nq * 32, 256), 256, 0, s>>>(S.stash.get(), S.splits.get(), S.nq, d_idx, d_d2);
        ASRB_CHECK_LAUNCH();
        DevBuf<int32_t> list((size_t)S.nq, s);
        DevBuf<int> count(1, s);
        ASRB_CUDA(cudaMemsetAsync(count.get(), 0, sizeof(int), s));
        heavy_list_kernel<<<grid_for((size_t)S.nq, 256), 256, 0, s>>>(S.splits.get(), S.nq, list.get(), count.get());
        ASRB_CHECK_LAUNCH();
        const unsigned blocks = (unsigned)std::min<size_t>((size_t)S.nq, 148 * 8);
        ball_query_heavy_kernel<<<blocks, 256, 0, s>>>(S.cells.view(), (const float4*)S.spts.get(), f, S.queries, S.radii,
                                                       S.splits.get(), list.get(), count.get(), keys.get(), d_idx, d_d2);
        ASRB_CHECK_LAUNCH();
    } else {
        ball_query_kernel<true><<<grid_for(S.nq, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                S.cells.view(), (const float4*)S.spts.get(), f, S.queries, S.radii, S.nq, nullptr, S.splits.get(), keys.get(),
                d_idx, d_d2, nullptr);
        ASRB_CHECK_LAUNCH();
    }
    S.stash.release();
}

void scale_compat(const float* d_sizes, const float* d_radii, const int32_t* d_idx, const int64_t* d_splits,
                  int64_t nq, float* d_out, cudaStream_t s) {
    if (nq == 0) return;
    scale_compat_kernel<<<grid_for((size_t)nq * 32, 256), 256, 0, s>>>(d_sizes, d_radii, d_idx, d_splits, nq, d_out);
    ASRB_CHECK_LAUNCH();
}



// =====================================================================================
// k-nearest-neighbour queries of the points among themselves — SURVEY.md §8 row f-2.
//
// Replaces KDTree::ComputeKRadius / ComputeInlier / ComputeRadiusNeighbors (reference
// cpp/lib/nsearch.cpp:30-105; nanoflann KD-tree + TBB).  Same Morton-sorted point array as the
// radius search; one warp owns a query point:
//   1. the 62 sorted neighbours of the point are loaded once and give, for every power-of-two
//      cell size, how many points share the point's cell -> the smallest cell size whose own
//      cell already holds k points is the starting scale;
//   2. the 3 x 3 x 3 block of cells around the point is streamed (windows by binary search in
//      the sorted codes), squared distances in nanoflann's operation order, and the k smallest
//      are kept as a sorted list spread over the lanes (insertion by ballot + shuffle);
//   3. the k-th distance is exact once it does not exceed the distance to the block's boundary
//      (>= one cell); otherwise the cell size doubles and the block is rescanned.
// k <= 64 (the reference's default is 24): one sorted list over the lanes for k <= 32, two for 32 < k <= 64.
namespace {

__device__ __forceinline__ long long lower_bound_code(const Key* __restrict__ a, long long n, Key k) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <bool INLIER, bool WIDE>  // WIDE: 32 < k <= 64, a second sorted list (ranks 32..63) over the lanes
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
knn_kernel(const Key* __restrict__ codes, const float4* __restrict__ spts, long long n, int k, float cell0,
           float* __restrict__ out_radius, const float* __restrict__ radii, float fraction, int vote_limit,
           uint8_t* __restrict__ out_inlier) {
    __shared__ unsigned s_begin[kWarpsPerBlock][32];
    __shared__ int s_pre[kWarpsPerBlock][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = blockIdx.x * (long long)kWarpsPerBlock + warp;
    if (i >= n) return;
    const float4 q = __ldg(spts + i);
    const Key code = __ldg(codes + i);
    // 1. starting scale
    const long long li = i - 1 - lane, ri = i + 1 + lane;
    const Key L = li >= 0 ? __ldg(codes + li) : ~Key(0), R = ri < n ? __ldg(codes + ri) : ~Key(0);
    int sh = 0;
    for (; sh < kGridBits; ++sh) {
        const Key pref = code >> (3 * sh);
        const unsigned ml = __ballot_sync(0xffffffffu, li >= 0 && (L >> (3 * sh)) == pref);
        const unsigned mr = __ballot_sync(0xffffffffu, ri < n && (R >> (3 * sh)) == pref);
        const int run = 1 + (__ffs(~ml) - 1 < 0 ? 32 : __ffs(~ml) - 1) + (__ffs(~mr) - 1 < 0 ? 32 : __ffs(~mr) - 1);
        if (run >= k) break;
    }
    float best_d = __int_as_float(0x7f800000);  // sorted ascending over the lanes
    int best_i = -1;
    float best_d1 = __int_as_float(0x7f800000);  // WIDE: ranks 32 .. 63
    int best_i1 = -1;
    auto kth_value = [&]() {
        return (WIDE && k > 32) ? __shfl_sync(0xffffffffu, best_d1, k - 33) : __shfl_sync(0xffffffffu, best_d, k - 1);
    };
    for (;;) {
        // 2. windows of the 27 cells around the point at this scale
        const int cells_per_axis = 1 << (kGridBits - sh);
        const Key cc = code >> (3 * sh);
        const int cx = (int)compact3(cc), cy = (int)compact3(cc >> 1), cz = (int)compact3(cc >> 2);
        unsigned begin = 0;
        int len = 0;
        if (lane < 27) {
            const int x = cx + lane % 3 - 1, y = cy + (lane / 3) % 3 - 1, z = cz + lane / 9 - 1;
            if (x >= 0 && y >= 0 && z >= 0 && x < cells_per_axis && y < cells_per_axis && z < cells_per_axis) {
                const Key c = morton3(x, y, z);
                const long long b = lower_bound_code(codes, n, c << (3 * sh));
                const long long e = sh == kGridBits ? n : lower_bound_code(codes, n, (c + 1) << (3 * sh));
                begin = (unsigned)b;
                len = (int)(e - b);
            }
        }
        int pre = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += v;
        }
        s_begin[warp][lane] = begin;
        s_pre[warp][lane + 1] = pre;
        if (lane == 0) s_pre[warp][0] = 0;
        __syncwarp();
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        best_d = best_d1 = __int_as_float(0x7f800000);
        best_i = best_i1 = -1;
        for (int t0 = 0; t0 < total; t0 += 32) {
            const int t = t0 + lane;
            float d2 = __int_as_float(0x7f800000);
            int pi = -1;
            if (t < total) {
                int c = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (c + step < 32 && s_pre[warp][c + step] <= t) c += step;
                const float4 p = __ldg(spts + s_begin[warp][c] + (unsigned)(t - s_pre[warp][c]));
                const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y), dz = __fsub_rn(q.z, p.z);
                d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                pi = __float_as_int(p.w);
            }
            const float kth = kth_value();
            unsigned acc = __ballot_sync(0xffffffffu, d2 < kth);
            while (acc) {
                const int src = __ffs(acc) - 1;
                acc &= acc - 1;
                const float c = __shfl_sync(0xffffffffu, d2, src);
                const int ci = __shfl_sync(0xffffffffu, pi, src);
                // later candidates of the batch may have fallen behind the updated k-th value
                if (!(c < kth_value())) continue;
                int pos = __popc(__ballot_sync(0xffffffffu, best_d <= c));
                const float up_d = __shfl_up_sync(0xffffffffu, best_d, 1);
                const int up_i = __shfl_up_sync(0xffffffffu, best_i, 1);
                if (WIDE) {
                    // the second list: shifted as a whole (rank 31 moves into its lane 0) or from the insertion point on
                    const int pos1 = __popc(__ballot_sync(0xffffffffu, best_d1 <= c));
                    const float up_d1 = __shfl_up_sync(0xffffffffu, best_d1, 1);
                    const int up_i1 = __shfl_up_sync(0xffffffffu, best_i1, 1);
                    const float last_d = __shfl_sync(0xffffffffu, best_d, 31);
                    const int last_i = __shfl_sync(0xffffffffu, best_i, 31);
                    if (pos < 32) {
                        best_d1 = lane == 0 ? last_d : up_d1;
                        best_i1 = lane == 0 ? last_i : up_i1;
                    } else {
                        if (lane > pos1) {
                            best_d1 = up_d1;
                            best_i1 = up_i1;
                        } else if (lane == pos1) {
                            best_d1 = c;
                            best_i1 = ci;
                        }
                        pos = 64;  // nothing changes in the first list
                    }
                }
                if (lane > pos) {
                    best_d = up_d;
                    best_i = up_i;
                } else if (lane == pos) {
                    best_d = c;
                    best_i = ci;
                }
            }
        }
        __syncwarp();
        // 3. exact once the k-th distance lies inside the scanned block
        const float kth = kth_value();
        const float cover = cell0 * (float)(1 << sh) * 0.9999f;
        if (sh >= kGridBits || kth <= cover * cover) break;
        ++sh;
    }
    const int valid0 = __popc(__ballot_sync(0xffffffffu, best_i >= 0 && lane < k));
    const int valid1 = WIDE ? __popc(__ballot_sync(0xffffffffu, best_i1 >= 0 && lane + 32 < k)) : 0;
    const int valid = valid0 + valid1;
    const long long self = __float_as_int(q.w);
    if (!INLIER) {
        float dmax = valid0 > 0 ? __shfl_sync(0xffffffffu, best_d, max(valid0, 1) - 1) : 0.f;
        if (WIDE) {
            const float d1 = __shfl_sync(0xffffffffu, best_d1, max(valid1, 1) - 1);
            if (valid1 > 0) dmax = d1;
        }
        if (lane == 0) out_radius[self] = sqrtf(dmax);
    } else {
        const float limit = __fmul_rn(radii[self], fraction);
        const bool vote = lane < valid0 && radii[best_i] < limit;
        int votes = __popc(__ballot_sync(0xffffffffu, vote));
        if (WIDE) {
            const bool vote1 = lane < valid1 && radii[max(best_i1, 0)] < limit;
            votes += __popc(__ballot_sync(0xffffffffu, vote1));
        }
        if (lane == 0) out_inlier[self] = votes < vote_limit ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
row_counts_kernel(const int64_t* __restrict__ splits, long long n, int32_t* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)(splits[i + 1] - splits[i]);
}
}  // namespace

void knn_build(Search& S, const float* d_points, int64_t n, cudaStream_t s) {
    // the sorted point array of the radius search, binned over the points' bounding cube
    S.n = n;
    S.nq = 0;
    DevBuf<unsigned> mm(6, s);
    ASRB_CUDA(cudaMemsetAsync(mm.get(), 0xff, 3 * sizeof(unsigned), s));
    ASRB_CUDA(cudaMemsetAsync(mm.get() + 3, 0, 3 * sizeof(unsigned), s));
    unsigned h[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0};
    if (n > 0) {
        const unsigned blocks = (unsigned)std::min<size_t>(grid_for(n, 256), 148 * 8);
        bbox_kernel<<<blocks, 256, 0, s>>>(d_points, n, mm.get(), mm.get() + 3);
        ASRB_CHECK_LAUNCH();
        ASRB_CUDA(cudaMemcpyAsync(h, mm.get(), sizeof(h), cudaMemcpyDeviceToHost, s));
        ASRB_CUDA(cudaStreamSynchronize(s));
    }
    float edge = 0.f;
    for (int a = 0; a < 3; ++a) {
        const float lo = h[a] == 0xffffffffu ? 0.f : from_ordered_bits(h[a]);
        const float hi = h[3 + a] == 0 ? 0.f : from_ordered_bits(h[3 + a]);
        S.frame_origin[a] = lo;
        edge = std::max(edge, hi - lo);
    }
    if (!(edge > 0.f)) edge = 1.f;
    S.frame_inv_h = (float)(2097152.0 / ((double)edge * 1.000001));
    BinFrame f{{S.frame_origin[0], S.frame_origin[1], S.frame_origin[2]}, S.frame_inv_h};
    S.codes.alloc((size_t)n, s);
    S.spts.alloc((size_t)n, s);
    DevBuf<uint32_t> order((size_t)n, s);
    if (n > 0) {
        ProfileScope prof("knn_point_sort", s);
        point_code_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_points, n, f, S.codes.get(), order.get());
        ASRB_CHECK_LAUNCH();
        sort_pairs_u64_u32(S.codes.get(), order.get(), (size_t)n, s, 63);
        gather_points_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_points, order.get(), n, (float4*)S.spts.get());
        ASRB_CHECK_LAUNCH();
    }
}

void knn_radius(const Search& S, int k, float* d_out, cudaStream_t s) {
    ASRB_REQUIRE(k >= 1 && k <= 64, "asr_b200 kNN: k must be in [1, 64] (the reference default is 24; the warp-wide candidate lists hold 64 entries)");
    if (S.n == 0) return;
    ProfileScope prof("knn_radius", s);
    if (k <= 32)
        knn_kernel<false, false><<<grid_for(S.n, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                S.codes.get(), (const float4*)S.spts.get(), S.n, k, 1.0f / S.frame_inv_h, d_out, nullptr, 0.f, 0, nullptr);
    else
        knn_kernel<false, true><<<grid_for(S.n, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                S.codes.get(), (const float4*)S.spts.get(), S.n, k, 1.0f / S.frame_inv_h, d_out, nullptr, 0.f, 0, nullptr);
    ASRB_CHECK_LAUNCH();
}

void knn_inlier(const Search& S, const float* d_radii, float fraction, int k, int outlier_threshold, uint8_t* d_out,
                cudaStream_t s) {
    ASRB_REQUIRE(k >= 1 && k <= 64, "asr_b200 kNN: k must be in [1, 64] (the reference default is 24; the warp-wide candidate lists hold 64 entries)");
    if (S.n == 0) return;
    ProfileScope prof("knn_inlier", s);
    if (k <= 32)
        knn_kernel<true, false><<<grid_for(S.n, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                S.codes.get(), (const float4*)S.spts.get(), S.n, k, 1.0f / S.frame_inv_h, nullptr, d_radii, fraction,
                outlier_threshold, d_out);
    else
        knn_kernel<true, true><<<grid_for(S.n, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                S.codes.get(), (const float4*)S.spts.get(), S.n, k, 1.0f / S.frame_inv_h, nullptr, d_radii, fraction,
                outlier_threshold, d_out);
    ASRB_CHECK_LAUNCH();
}

// number of points with |p - p_i|^2 < r_i^2 for every point (ComputeRadiusNeighbors, nsearch.cpp:87-105)
void radius_neighbor_counts(const float* d_points, int64_t n, const float* d_radii, int32_t* d_out, cudaStream_t s) {
    if (n == 0) return;
    Search S;
    S.want_fill = false;  // counts only
    search_prepare(S, d_points, n, d_points, d_radii, n, nullptr, s);
    row_counts_kernel<<<grid_for(n, 256), 256, 0, s>>>(S.splits.get(), n, d_out);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
