#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <vector>

namespace ttb {

namespace {
thread_local std::string g_last_error;

struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
};
std::vector<ProfRec> g_prof;
std::vector<size_t> g_prof_open;  // indices of the scopes that are open (scopes may nest)
std::vector<cudaEvent_t> g_free_events;
cudaEvent_t take_event() {
    if (!g_free_events.empty()) {
        cudaEvent_t e = g_free_events.back();
        g_free_events.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

bool prof_enabled() {
    static const bool on = getenv("TTB_PROF") != nullptr;
    return on;
}
void prof_begin(const char* name, cudaStream_t stream) {
    ProfRec r{name, take_event(), take_event()};
    cudaEventRecord(r.e0, stream);
    g_prof_open.push_back(g_prof.size());
    g_prof.push_back(r);
}
void prof_end(cudaStream_t stream) {
    if (g_prof_open.empty()) return;
    const size_t idx = g_prof_open.back();
    g_prof_open.pop_back();
    if (idx < g_prof.size()) cudaEventRecord(g_prof[idx].e1, stream);
}
void prof_report(const char* title) {
    if (!prof_enabled() || g_prof.empty()) return;
    cudaEventSynchronize(g_prof.back().e1);
    std::map<std::string, std::pair<double, int>> agg;
    double total = 0.0;
    for (const ProfRec& r : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        agg[r.name].first += ms;
        agg[r.name].second += 1;
        total += ms;
        g_free_events.push_back(r.e0);
        g_free_events.push_back(r.e1);
    }
    g_prof.clear();
    g_prof_open.clear();
    std::vector<std::pair<std::string, std::pair<double, int>>> v(agg.begin(), agg.end());
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.second.first > b.second.first; });
    fprintf(stderr, "[prof] %s: %.3f ms in scopes\n", title, total);
    for (const auto& kv : v)
        fprintf(stderr, "[prof]   %-28s n=%5d total=%9.3f ms avg=%8.2f us\n", kv.first.c_str(), kv.second.second,
                kv.second.first, 1e3 * kv.second.first / kv.second.second);
}

}  // namespace ttb
