// Optional per-kernel timing with CUDA events on the launching stream.
// Disabled by default (zero overhead beyond one branch); bench.py enables it to
// measure the dominant kernel's average launch duration inside the timed region.
#pragma once
#include <cstdlib>
#include <ctime>

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

// ASR_DEBUG_TIMING=1: host-synchronised wall time of the phases of a build on stderr (development aid)
struct PhaseTimer {
    bool on;
    cudaStream_t s;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit PhaseTimer(cudaStream_t stream) : on(getenv("ASR_DEBUG_TIMING") != nullptr), s(stream), t0(0) {
        if (on) {
            cudaStreamSynchronize(s);
            t0 = now();
        }
    }
    void lap(const char* what, long long n = -1) {
        if (!on) return;
        cudaStreamSynchronize(s);
        const double t = now();
        fprintf(stderr, "[asr timing] %-28s %8.3f ms  (n = %lld)\n", what, t - t0, n);
        t0 = t;
    }
};

}  // namespace asrb
