// Open-addressing hash table  location code -> position  (linear probing,
// 16-byte entries so a probe is one aligned 128-bit load; load factor <= 0.5).
// Replaces the std::lower_bound / libcuckoo lookups of the reference's grid code
// (grid.cpp:85-92, octreebase.h:174-182) with ~1.5 probes instead of ~22 dependent
// binary-search steps.
#pragma once
#include "common.cuh"

namespace asrb {

struct HashEntry {
    Key key;
    long long val;
};

struct KeyTableView {
    const HashEntry* e;
    uint32_t mask;
};

__host__ __device__ __forceinline__ uint32_t hash_key(Key k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (uint32_t)k;
}

// position of `k`, or -1
__device__ __forceinline__ long long table_find(const KeyTableView t, Key k) {
    uint32_t s = hash_key(k) & t.mask;
    for (;;) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(t.e + s));
        if (v.x == k) return (long long)v.y;
        if (v.x == kNoKey) return -1;
        s = (s + 1) & t.mask;
    }
}

struct KeyTable {
    DevBuf<HashEntry> entries;
    uint32_t mask = 0;
    KeyTableView view() const { return KeyTableView{entries.get(), mask}; }
    // keys must be unique and != kNoKey; value of keys[i] is i; `reserve`: capacity for that many keys in total
    void build(const Key* d_keys, size_t n, cudaStream_t s, size_t reserve = 0);
    // adds keys (value base + i) to a table whose capacity allows it (load factor stays <= 0.5 up to `reserve`)
    void insert(const Key* d_keys, size_t n, size_t base, cudaStream_t s);
    size_t capacity() const { return (size_t)mask + 1; }
};

}  // namespace asrb
