// "gx" sparse convolution: the U-Net's generalized sparse convolutions (reference
// models/common_torch.py:95-148 -> Open3D sparse_conv) for activations kept in the
// split-half format between layers.  See spconv_gx.cu.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace asrb {
namespace gx {

constexpr int kTM = 128;  // rows of one MMA tile (TMEM lanes)

// A [V, C] activation tensor stored as two fp16 planes inside rows of `pitch` halves: x = hi + lo,
// hi at columns [hi, hi + C), lo at [lo, lo + C).  A plain tensor has pitch = 2 C, hi = 0, lo = C; a
// channel slice of a concatenation buffer keeps the buffer's pitch.  Row `rows - 1` is all zero (the
// gather target of absent neighbours), so buffers hold V + 1 rows.
struct H2View {
    __half* p = nullptr;
    int64_t rows = 0;  // V + 1
    int C = 0, pitch = 0, hi = 0, lo = 0;
};

enum PlanMode {
    kModeStationary = 0,  // dense slots output-stationary + rare slots through the pair buffer
    kModePairFinal = 1,   // every row has exactly one entry (up tables): pair-major, final epilogue
};

struct Plan {
    int mode = kModeStationary;
    int64_t V = 0, V_in = 0, E = 0;  // output rows, input rows (gather index V_in = the zero row), entries
    int K = 0, D = 0;          // D = number of dense slots (slots 0 .. D-1)
    int64_t T = 0;             // output-stationary tiles = ceil(V / 128)
    DevBuf<int32_t> gidx;      // [T][D][128] gather rows of the dense slots (zero row where absent)
    DevBuf<int64_t> rare_rs;   // [V + 1] rare entries before each row (row-major rare order)
    DevBuf<int32_t> rare_in;   // [R] input row of every rare entry, same order
    int64_t R = -1;            // number of rare entries (known on the host after finish())
    int64_t* R_host = nullptr; // pinned
    cudaEvent_t R_event = nullptr;
    // pair-major tiles of the rare entries (or of all entries in kModePairFinal)
    int max_pair_tiles = 0;
    DevBuf<int32_t> pt_slot;   // [tiles] kernel slot
    DevBuf<int32_t> pt_gidx;   // [tiles][128] gather rows
    DevBuf<int32_t> pt_out;    // [tiles][128] pair-buffer row / output row, -1 = padding
    DevBuf<int> pt_count;      // device scalar: number of pair tiles
    // a plan may be built on one stream and used on another (after the caller's event wait): its buffers are
    // then FREED in the order of the stream that uses them
    void rehome(cudaStream_t s) {
        gidx.s = rare_rs.s = rare_in.s = pt_slot.s = pt_gidx.s = pt_out.s = pt_count.s = s;
    }
    // begin() state kept for finish()
    const int32_t* d_idx = nullptr;
    const uint8_t* d_slot = nullptr;
    const int64_t* d_splits = nullptr;
    // the plan covers the table rows d_row_map[0 .. V) only (null: all rows, local row = table row); outputs,
    // normalisers and residuals stay indexed by table row.  The caller keeps the array alive with the plan.
    const int32_t* d_row_map = nullptr;
    bool finished = false;
    ~Plan();
};

// Two-phase build so that several tables share ONE host synchronisation: begin() queues the counting
// kernels and an async copy of the rare-entry count; finish() waits for it and builds the tile lists.
void plan_begin(Plan& P, const int32_t* d_idx, const uint8_t* d_slot, const int64_t* d_splits, int64_t V, int64_t V_in,
                int64_t E, int K, int mode, const int32_t* d_row_map, cudaStream_t s);
void plan_finish(Plan& P, cudaStream_t s);

// Packed filter bank for the tensor-core kernel: per slot and 64-channel chunk the fp16 hi / lo parts
// of W * 2^scale_exp for output columns [col0, col0 + ncols), N = ncols rounded up to 16, in the shared-
// memory image the kernel copies in one piece (K-major, 128-byte swizzle).
size_t packed_filter_bytes(int K, int Cin, int ncols);
void pack_filters(const float* W, int K, int Cin, int Cout, int col0, int ncols, int scale_exp, void* out,
                  cudaStream_t s);

struct ConvArgs {
    H2View x;                 // input activations (x.C = Cin)
    const void* wp = nullptr; // packed filters
    int K = 0, ncols = 0;     // real output columns of this call
    int scale_exp = 0;        // the packed filters hold W * 2^scale_exp
    const float* bias = nullptr;   // [ncols] or null
    int relu = 0;
    const float* norm = nullptr;   // [V] divide every column by norm[row] where != 0 (before the bias), or null
    const float* imp = nullptr;    // [>= V_in] importance of every INPUT row: out = sum_n imp[idx_n] x[idx_n] W[slot_n]
                                   // (conv1b of SpecialSparseConv, common_torch.py:124-142); <= 128 columns
    H2View res;               // optional residual added AFTER the activation (res.p == null: none)
    H2View out;               // h2 output view (out.p != null) ...
    float* out_f32 = nullptr; // ... or fp32 output [V, out_f32_pitch] at column out_f32_col
    int out_f32_pitch = 0, out_f32_col = 0;
    float* pairbuf = nullptr; // scratch for the rare entries: [R][N] floats (N = ncols rounded up to 16)
};
size_t pairbuf_floats(const Plan& P, int ncols);
void set_acc_groups(int g);
void set_tma_gather(int v);
void set_l1_gather(int v);
void set_max_stages(int v);
void set_single_tmem(int v);
void set_one_team(int v);
void set_trace(int v);
// dev: 16 unsigned counters per CTA (<= 256 CTAs) of the last traced launch, see spconv_gx.cu
void trace_read(unsigned* host, int ctas, cudaStream_t s);
void set_ablate(int v);  // timing experiments only
void conv(const Plan& P, const ConvArgs& a, cudaStream_t s);

// format conversion and the elementwise helpers of the split-half format
// rows (may be null): input row i is written to output row rows[i] (only those rows of `out` are touched)
void from_f32(const float* x, int64_t V, int C, int ldx, const float* row_scale, const int32_t* rows, H2View out,
              cudaStream_t s);
void to_f32(H2View x, int64_t V, float* out, int ldo, cudaStream_t s);
void scale_rows(H2View x, int64_t V, const float* row_scale, H2View out, cudaStream_t s);
int overflow_flag_read_and_clear(cudaStream_t s);  // 1 if any conversion saturated since the last call (synchronises)

}  // namespace gx
}  // namespace asrb
