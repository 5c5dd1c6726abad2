#include "hash.cuh"
#include "profile.cuh"

namespace asrb {

__global__ void __launch_bounds__(256)
table_clear_kernel(HashEntry* e, size_t cap) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < cap) {
        e[i].key = kNoKey;
        e[i].val = -1;
    }
}

__global__ void __launch_bounds__(256)
table_insert_kernel(const Key* __restrict__ keys, size_t n, HashEntry* e, uint32_t mask, long long base) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key k = keys[i];
    uint32_t s = hash_key(k) & mask;
    for (;;) {
        const Key prev = atomicCAS(&e[s].key, kNoKey, k);
        if (prev == kNoKey || prev == k) {
            e[s].val = base + (long long)i;
            return;
        }
        s = (s + 1) & mask;
    }
}

// membership-only insertion of more keys into a table built with spare capacity (values = base + i)
void KeyTable::insert(const Key* d_keys, size_t n, size_t base, cudaStream_t s) {
    if (!n) return;
    table_insert_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_keys, n, entries.get(), mask, (long long)base);
    ASRB_CHECK_LAUNCH();
}

void KeyTable::build(const Key* d_keys, size_t n, cudaStream_t s, size_t reserve) {
    size_t cap = 64;
    while (cap < 2 * std::max(n, reserve)) cap <<= 1;
    mask = (uint32_t)(cap - 1);
    entries.alloc(cap, s);
    ProfileScope prof("hash_build", s);
    table_clear_kernel<<<grid_for(cap, 256), 256, 0, s>>>(entries.get(), cap);
    ASRB_CHECK_LAUNCH();
    if (n) {
        table_insert_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_keys, n, entries.get(), mask, 0);
        ASRB_CHECK_LAUNCH();
    }
}

}  // namespace asrb
