// Per-launch device timing for bench.py's roofline report: when enabled, every BARK_LAUNCH is bracketed by CUDA
// events on the launching stream; bark_b200_profile_report() aggregates elapsed time and annotated algorithmic
// work per kernel name.  Off by default (events add ~2 us of host time per launch).
#include "../../include/bark_b200.h"
#include "common.cuh"

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace bark {

std::atomic<unsigned long long> g_h2d_bytes{0}, g_d2h_bytes{0};
bool g_prof_on = false;
thread_local double g_next_bytes = 0.0, g_next_flops = 0.0;

namespace {
struct Rec { const char * name; cudaEvent_t a, b; double bytes, flops; };
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
std::mutex g_prof_mutex;                           // profiling is a single-context measurement tool; the lock only keeps the lists consistent
cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; BARK_CUDA_CHECK(cudaEventCreate(&e)); return e;
}
}  // namespace

void prof_begin(const char * name, cudaStream_t s, double bytes, double flops) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    Rec r{name, get_event(), get_event(), bytes, flops};
    BARK_CUDA_CHECK(cudaEventRecord(r.a, s));
    g_recs.push_back(r);
}
void prof_end(cudaStream_t s) { std::lock_guard<std::mutex> lock(g_prof_mutex); BARK_CUDA_CHECK(cudaEventRecord(g_recs.back().b, s)); }

}  // namespace bark

using namespace bark;

extern "C" void bark_b200_profile_enable(int on) {
    g_prof_on = on != 0;
    for (auto & r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
    g_recs.clear();
}

// JSON: {"kernel": {"launches": n, "ms": t, "bytes": b, "flops": f}, ...}; returns the length needed
static int bark_b200_profile_report_impl(char * buf, int cap) {
    BARK_CUDA_CHECK(cudaDeviceSynchronize());
    struct Agg { long n = 0; double ms = 0, bytes = 0, flops = 0; };
    std::map<std::string, Agg> agg;
    for (auto & r : g_recs) {
        float ms = 0; BARK_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
        std::string n = r.name;
        if (!n.empty() && n.front() == '(') n = n.substr(1, n.size() - 2);
        Agg & a = agg[n]; a.n++; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
    }
    std::string out = "{";
    bool first = true;
    for (auto & kv : agg) {
        char tmp[512];
        snprintf(tmp, sizeof tmp, "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f, \"bytes\": %.6e, \"flops\": %.6e}", first ? "" : ", ", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes, kv.second.flops);
        out += tmp; first = false;
    }
    out += "}";
    if (buf && cap > 0) { snprintf(buf, (size_t) cap, "%s", out.c_str()); }
    return (int) out.size() + 1;
}
extern "C" int bark_b200_profile_report(char * buf, int cap) { return guarded((int) 0, [&] { return bark_b200_profile_report_impl(buf, cap); }); }

extern "C" void bark_b200_io_counters(unsigned long long * h2d, unsigned long long * d2h, int reset) {
    if (h2d) *h2d = g_h2d_bytes.load();
    if (d2h) *d2h = g_d2h_bytes.load();
    if (reset) { g_h2d_bytes = 0; g_d2h_bytes = 0; }
}
