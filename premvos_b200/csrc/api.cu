// Library-wide C ABI helpers: version string, thread-local error text, launch counter.
#include <stdarg.h>

#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace premvos {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launch_count{0};

void set_last_error(const std::string& msg) { t_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

// ---- per-launch profiling ----------------------------------------------------------------------
struct LaunchRec { const char* name; double flops, bytes; cudaEvent_t e0, e1; };
static std::atomic<bool> g_profiling{false};
static std::vector<LaunchRec> g_recs;
static std::mutex g_prof_mutex;                            // launchers may be called from several host threads
static thread_local cudaEvent_t t_pending_e0 = nullptr;    // the start event of the launch this thread is bracketing

bool profiling_enabled() { return g_profiling; }
// stable storage for dynamically built profile labels (per-layer breakdowns, PREMVOS_PROFILE_LAYERS=1)
const char* prof_intern(const std::string& s) {
  static std::map<std::string, int> pool;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  return pool.emplace(s, 0).first->first.c_str();
}
void prof_before(cudaStream_t st) {
  if (!g_profiling) return;
  if (t_pending_e0) cudaEventDestroy(t_pending_e0);        // a launch that failed between before / after: drop its event
  cudaEventCreate(&t_pending_e0);
  cudaEventRecord(t_pending_e0, st);
}
void prof_after(const char* what, cudaStream_t st, double flops, double bytes) {
  LaunchRec r;
  r.name = what; r.flops = flops; r.bytes = bytes; r.e0 = t_pending_e0;
  t_pending_e0 = nullptr;
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e1, st);
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_recs.push_back(r);
}

}  // namespace premvos

extern "C" int premvos_profile_begin(void) {
  std::lock_guard<std::mutex> lock(premvos::g_prof_mutex);
  premvos::g_recs.clear();
  premvos::g_profiling = true;
  return 0;
}

// Stops profiling and writes one line per kernel name: "name count total_ms flops bytes\n".  Returns 0, or -- when `buflen` is too
// small for the report -- the number of bytes needed (nothing is written then; call again with a larger buffer: the report is kept).
extern "C" int premvos_profile_end(char* buf, int buflen) {
  using namespace premvos;
  static std::string kept;       // a report that did not fit the caller's buffer
  g_profiling = false;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail((int)e, "premvos_profile_end: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (!kept.empty() && g_recs.empty()) {
    if (!buf || buflen <= (int)kept.size()) return (int)kept.size() + 1;
    memcpy(buf, kept.c_str(), kept.size() + 1);
    kept.clear();
    return 0;
  }
  struct Agg { int count = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (r.e0 && r.e1) cudaEventElapsedTime(&ms, r.e0, r.e1);
    Agg& a = agg[r.name];
    a.count++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    if (r.e0) cudaEventDestroy(r.e0);
    if (r.e1) cudaEventDestroy(r.e1);
  }
  g_recs.clear();
  std::string out;
  char line[512];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %d %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.count, kv.second.ms,
             kv.second.flops, kv.second.bytes);
    out += line;
  }
  if (!buf || buflen <= (int)out.size()) {
    kept = out;
    return (int)out.size() + 1;
  }
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

extern "C" const char* premvos_version(void) { return "premvos_b200 0.1 sm_100a"; }
extern "C" const char* premvos_last_error(void) { return premvos::t_last_error.c_str(); }
extern "C" int64_t premvos_kernel_launch_count(void) { return premvos::g_launch_count.load(); }
