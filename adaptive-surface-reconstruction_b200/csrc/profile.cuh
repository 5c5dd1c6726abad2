// Optional per-kernel timing with CUDA events on the launching stream.
// Disabled by default (zero overhead beyond one branch); bench.py enables it to
// measure the dominant kernel's average launch duration inside the timed region.
#pragma once
#include "common.cuh"

namespace asrb {

bool profile_enabled();
void profile_push(const char* name, cudaEvent_t start, cudaEvent_t stop, double flops);
cudaEvent_t profile_event();

struct ProfileScope {
    const char* name;
    cudaStream_t s;
    cudaEvent_t e0 = nullptr;
    double flops;
    ProfileScope(const char* n, cudaStream_t stream, double fl = 0.0) : name(n), s(stream), flops(fl) {
        if (profile_enabled()) {
            e0 = profile_event();
            cudaEventRecord(e0, s);
        }
    }
    ~ProfileScope() {
        if (e0) {
            cudaEvent_t e1 = profile_event();
            cudaEventRecord(e1, s);
            profile_push(name, e0, e1, flops);
        }
    }
};

}  // namespace asrb
