// Thin wrappers over CUB device-wide primitives (sort / scan / select) used as
// plumbing between the hand-written kernels.  All calls are asynchronous on
// `s`; temp storage comes from the stream-ordered pool.
#pragma once
#include "common.cuh"

namespace asrb {

// ascending radix sort of 64-bit keys, in place (bits [0, end_bit))
void sort_keys_u64(Key* d_keys, size_t n, cudaStream_t s, int end_bit = 64);
// ascending stable sort of (key, value) pairs, in place
void sort_pairs_u64_u32(Key* d_keys, uint32_t* d_vals, size_t n, cudaStream_t s, int end_bit = 64);
void sort_pairs_u64_u64(Key* d_keys, unsigned long long* d_vals, size_t n, cudaStream_t s, int end_bit = 64);
void sort_pairs_u32_u32(uint32_t* d_keys, uint32_t* d_vals, size_t n, cudaStream_t s, int end_bit = 32);
void sort_pairs_u8_u32(uint8_t* d_keys, uint32_t* d_vals, size_t n, cudaStream_t s, int end_bit = 8);
// removes consecutive duplicates of a sorted array in place; returns new length (synchronises)
size_t unique_u64(Key* d_keys, size_t n, cudaStream_t s);
// out[i] = sum_{j<i} in[j]; out has n+1 entries (out[n] = total)
void exclusive_sum_i32_to_i64(const int32_t* d_in, int64_t* d_out, size_t n, cudaStream_t s);
void exclusive_sum_u8_to_i64(const uint8_t* d_in, int64_t* d_out, size_t n, cudaStream_t s);

// pooled pinned host slots for asynchronous scalar read-backs
int64_t* pinned_slot_acquire();
void pinned_slot_release(int64_t* p);

}  // namespace asrb
