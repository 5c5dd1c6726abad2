// Multi-GPU helpers of the sharded hot path (one process per GPU, geometry replicated, rows of every grid level
// owned by Z-curve region).  All device-side, no host synchronisation:
//
//   shard_positions   position of every voxel on the Z-curve at depth 21 (any octree level)
//   shard_owner       rank that owns a voxel = number of region thresholds <= its position
//   shard_need_mask   for a neighbour table: which ranks read each row that THIS rank owns (bit r = rank r)
//   shard_push        copy the owned rows that other ranks read into the SAME buffer on those ranks, through
//                     peer-mapped (NVLink / NVSwitch) pointers of a symmetric allocation — the halo exchange
//                     is plain stores from this GPU, no collective call and no packing
#include "internal.h"
#include "profile.cuh"

namespace asrb {

namespace {
__device__ __forceinline__ unsigned long long zcurve_position(Key k) {
    const int lev = key_level(k);
    return (k ^ (Key(1) << (3 * lev))) << (3 * (kMaxLevel - lev));
}

__global__ void __launch_bounds__(256)
shard_positions_kernel(const Key* __restrict__ keys, long long V, unsigned long long* __restrict__ pos) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < V) pos[i] = zcurve_position(keys[i]);
}

__global__ void __launch_bounds__(256)
shard_owner_kernel(const Key* __restrict__ keys, long long V, const unsigned long long* __restrict__ thr, int nthr,
                   uint8_t* __restrict__ owner) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V) return;
    const unsigned long long p = zcurve_position(keys[i]);
    int o = 0;
    for (int t = 0; t < nthr; ++t) o += thr[t] <= p;
    owner[i] = (uint8_t)o;
}

// thread per output row of the table: entries whose input row is mine and whose output row is somebody else's
__global__ void __launch_bounds__(256)
shard_need_mask_kernel(const int64_t* __restrict__ splits, const int32_t* __restrict__ idx, long long V_out,
                       const uint8_t* __restrict__ owner_out, const uint8_t* __restrict__ owner_in, int me,
                       uint32_t* __restrict__ mask) {
    const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (o >= V_out) return;
    const int oo = owner_out[o];
    if (oo == me) return;
    for (int64_t j = splits[o]; j < splits[o + 1]; ++j) {
        const int n = idx[j];
        if (owner_in[n] == me) atomicOr(mask + n, 1u << oo);
    }
}

struct PushArgs {
    unsigned char* peer[8];  // base of the symmetric allocation on every rank (peer-mapped), [me] unused
    long long offset;        // byte offset of the buffer inside the allocation
    long long pitch;         // bytes per row
    int seg_off[2];          // byte offsets of the (up to two) row segments to copy
    int seg_len;             // bytes per segment (multiple of 16)
    int nseg;
    const int32_t* rows;     // rows this rank owns
    long long nrows;
    const uint32_t* mask;    // ranks that read each row (null: all_mask for every row)
    uint32_t all_mask;
    int me;
};

// 8 lanes per row: 16-byte pieces, one row segment per pass
__global__ void __launch_bounds__(256) shard_push_kernel(PushArgs a) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long r = t >> 3;
    const int sub = threadIdx.x & 7;
    if (r >= a.nrows) return;
    const long long row = a.rows[r];
    uint32_t m = a.mask ? a.mask[row] : a.all_mask;
    m &= ~(1u << a.me);
    if (!m) return;
    const unsigned char* src = a.peer[a.me] + a.offset + row * a.pitch;
    for (int sgi = 0; sgi < a.nseg; ++sgi) {
        for (int b = sub * 16; b < a.seg_len; b += 128) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + a.seg_off[sgi] + b);
            uint32_t mm = m;
            while (mm) {
                const int p = __ffs(mm) - 1;
                mm &= mm - 1;
                *reinterpret_cast<uint4*>(a.peer[p] + a.offset + row * a.pitch + a.seg_off[sgi] + b) = v;
            }
        }
    }
}
}  // namespace

void shard_positions(const Key* keys, int64_t V, unsigned long long* pos, cudaStream_t s) {
    if (V == 0) return;
    shard_positions_kernel<<<grid_for((size_t)V, 256), 256, 0, s>>>(keys, V, pos);
    ASRB_CHECK_LAUNCH();
}

void shard_owner(const Key* keys, int64_t V, const unsigned long long* thr, int nthr, uint8_t* owner, cudaStream_t s) {
    ASRB_REQUIRE(nthr >= 0 && nthr < 8, "shard_owner: at most 8 ranks");
    if (V == 0) return;
    shard_owner_kernel<<<grid_for((size_t)V, 256), 256, 0, s>>>(keys, V, thr, nthr, owner);
    ASRB_CHECK_LAUNCH();
}

void shard_need_mask(const int64_t* splits, const int32_t* idx, int64_t V_out, const uint8_t* owner_out,
                     const uint8_t* owner_in, int me, uint32_t* mask, cudaStream_t s) {
    if (V_out == 0) return;
    ProfileScope prof("shard_need_mask", s);
    shard_need_mask_kernel<<<grid_for((size_t)V_out, 256), 256, 0, s>>>(splits, idx, V_out, owner_out, owner_in, me, mask);
    ASRB_CHECK_LAUNCH();
}

void shard_push(void* const* peer_base, int world, int me, int64_t offset, int64_t pitch, const int* seg_off, int nseg,
                int seg_len, const int32_t* rows, int64_t nrows, const uint32_t* mask, cudaStream_t s) {
    ASRB_REQUIRE(world >= 1 && world <= 8 && me >= 0 && me < world, "shard_push: bad rank / world size");
    ASRB_REQUIRE(nseg >= 1 && nseg <= 2 && seg_len > 0 && seg_len % 16 == 0 && pitch % 16 == 0 && offset % 16 == 0,
                 "shard_push: segments must be 16-byte aligned");
    if (nrows == 0 || world == 1) return;
    PushArgs a{};
    for (int r = 0; r < world; ++r) a.peer[r] = (unsigned char*)peer_base[r];
    a.offset = offset;
    a.pitch = pitch;
    for (int i = 0; i < nseg; ++i) {
        ASRB_REQUIRE(seg_off[i] % 16 == 0, "shard_push: segments must be 16-byte aligned");
        a.seg_off[i] = seg_off[i];
    }
    a.seg_len = seg_len;
    a.nseg = nseg;
    a.rows = rows;
    a.nrows = nrows;
    a.mask = mask;
    a.all_mask = (1u << world) - 1u;
    a.me = me;
    ProfileScope prof("shard_push", s);
    shard_push_kernel<<<grid_for((size_t)nrows * 8, 256), 256, 0, s>>>(a);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
