#pragma once
#include <memory>

#include "common.cuh"

namespace asrb {

struct Int4Pod {
    int x, y, z, w;
};

// Slot-sorted view of one neighbour table; reused by every convolution on it.
struct ConvPlan {
    int64_t V_out = 0, E = 0;
    int K = 0;
    int max_tiles = 0;
    DevBuf<int32_t> p_in, p_out;  // [E] pairs sorted (stably) by (row block, kernel slot)
    DevBuf<uint32_t> perm;        // [E] original CSR position of each sorted pair
    DevBuf<Int4Pod> tiles;        // (slot, first pair, count, -)
    DevBuf<int> num_tiles;        // device scalar
    // (row block, slot) groups, g = block * K + slot: first pair and number of 128-pair tiles
    // before each group ([G + 1] each); the persistent kernel walks these instead of `tiles`
    int G = 0, num_blocks = 1;
    // > 0: the first tiles0 tiles are the slot-0 tiles and cover every output row exactly once
    mutable int tiles0 = 0;
    int tiles0_if_flag = 0;
    mutable bool flag_pending = false;  // the device flag is still on its way to *flag_host
    int* flag_host = nullptr;
    cudaEvent_t flag_event = nullptr;
    ~ConvPlan();
    DevBuf<long long> g_begin;
    DevBuf<int> g_tile0;
    bool has_tiles2 = false;      // built only when the 2-row-group option is on at plan creation
    int max_tiles2 = 0;           // same for 256-pair tiles (tensor-core kernel, two 128-row MMA groups)
    DevBuf<Int4Pod> tiles2;
    DevBuf<int> num_tiles2;
};

int* flag_slot_acquire();
void flag_slot_release(int* p);
inline ConvPlan::~ConvPlan() {
    if (flag_event) {
        if (flag_pending) cudaEventSynchronize(flag_event);  // the async copy into the slot must have landed
        cudaEventDestroy(flag_event);
    }
    flag_slot_release(flag_host);
}

void conv_plan_build(ConvPlan& P, const int32_t* d_idx, const uint8_t* d_slot, const int64_t* d_splits, int64_t V_out,
                     int64_t E, int K, cudaStream_t s);

// out must hold V_out*Cout floats.  imp_in (per input row, gathered through the
// index) and/or imp_entry (per CSR entry) weight channels >= imp_col;
// normalize divides channels >= norm_col by norm[row] (or the row length when
// norm is null) where that is non-zero.
// wp != nullptr selects the tensor-core tile kernel (filters packed by pack_conv_filters).
void sparse_conv_forward(const ConvPlan& P, const float* x, const float* w, const float* wp, int Cin, int Cout,
                         const float* imp_in,
                         const float* imp_entry, int imp_col, int normalize, int norm_col, const float* norm,
                         const int64_t* splits, const float* bias, int relu, float* out, cudaStream_t s);

size_t packed_conv_filters_floats(int K, int Cin, int Cout);
void pack_conv_filters(const float* W, int K, int Cin, int Cout, float* out, cudaStream_t s);
// store_first: `out` is NOT zero-initialised; the slot-0 tiles (P.tiles0 > 0) write it first
void sparse_conv_tc_tiles(const ConvPlan& P, const float* x, const float* wp, int Cin, int Cout, const float* imp_in,
                          const float* imp_entry, int imp_col, float* out, cudaStream_t s, bool store_first = false);

void sparse_conv_tc_tune(int stages, int mt);
void sparse_conv_row_block_shift(int v);
void sparse_conv_tc_ntile(int n);
int sparse_conv_tc_row_groups();

void row_importance(const float* imp, const int32_t* idx, const int64_t* splits, int64_t V, float* out,
                    cudaStream_t s);

}  // namespace asrb
