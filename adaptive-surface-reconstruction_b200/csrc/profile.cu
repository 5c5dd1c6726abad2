#include "profile.cuh"

#include <map>
#include <mutex>
#include <vector>

namespace asrb {

namespace {
struct Rec {
    cudaEvent_t a, b;
    double flops;
};
struct Acc {
    double ms = 0, flops = 0;
    long long launches = 0;
};
std::mutex g_mu;
bool g_on = false;
std::map<std::string, std::vector<Rec>> g_pending;
std::map<std::string, Acc> g_acc;
std::vector<cudaEvent_t> g_free;

void drain_locked() {
    for (auto& kv : g_pending) {
        Acc& a = g_acc[kv.first];
        for (Rec& r : kv.second) {
            cudaEventSynchronize(r.b);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) a.ms += ms;
            a.flops += r.flops;
            a.launches += 1;
            g_free.push_back(r.a);
            g_free.push_back(r.b);
        }
        kv.second.clear();
    }
}
}  // namespace

bool profile_enabled() { return g_on; }

cudaEvent_t profile_event() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_free.empty()) {
        cudaEvent_t e = g_free.back();
        g_free.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void profile_push(const char* name, cudaEvent_t a, cudaEvent_t b, double flops) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_pending[name].push_back(Rec{a, b, flops});
}

void profile_set(bool on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_on = on;
}
void profile_reset() {
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    g_acc.clear();
}
int profile_count() {
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    return (int)g_acc.size();
}
bool profile_get(int i, std::string& name, double& ms, long long& launches, double& flops) {
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    if (i < 0 || i >= (int)g_acc.size()) return false;
    auto it = g_acc.begin();
    std::advance(it, i);
    name = it->first;
    ms = it->second.ms;
    launches = it->second.launches;
    flops = it->second.flops;
    return true;
}

}  // namespace asrb
