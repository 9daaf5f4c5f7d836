// prof.cu -- opt-in per-kernel device timing (CUDA events on the launching stream) for bench.py's roofline
// object: the dominant kernels are launched from inside multi-kernel C-ABI calls, so the caller cannot bracket
// them with its own events.  Off by default; must not be enabled during CUDA-graph capture.
#include <mutex>
#include <vector>
#include "common.cuh"

namespace yolat {

namespace {
struct Slot { cudaEvent_t a, b; int id; };
std::mutex g_mu;
bool g_on = false;
std::vector<Slot> g_pool;      // recorded event pairs since the last reset
std::vector<Slot> g_free;
constexpr size_t kMaxPairs = 8192;
}  // namespace

ProfScope::ProfScope(int id, cudaStream_t st) : st_(st), active_(false), idx_(0) {
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_on || g_pool.size() >= kMaxPairs) return;
  Slot s;
  if (!g_free.empty()) {
    s = g_free.back();
    g_free.pop_back();
  } else {
    if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
  }
  s.id = id;
  cudaEventRecord(s.a, st_);
  g_pool.push_back(s);
  idx_ = g_pool.size() - 1;
  active_ = true;
}

ProfScope::~ProfScope() {
  if (!active_) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (idx_ < g_pool.size()) cudaEventRecord(g_pool[idx_].b, st_);
}

}  // namespace yolat

using namespace yolat;

extern "C" {

int yolat_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  if (g_on) {   // start a fresh measurement window
    for (auto& s : g_pool) g_free.push_back(s);
    g_pool.clear();
  }
  return YOLAT_OK;
}

int yolat_prof_read(int id, int64_t* launches, double* total_ms) {
  if (!launches || !total_ms) return YOLAT_ERR_INVALID;
  std::lock_guard<std::mutex> lk(g_mu);
  *launches = 0;
  *total_ms = 0.0;
  for (auto& s : g_pool) {
    if (s.id != id) continue;
    if (cudaEventSynchronize(s.b) != cudaSuccess) return YOLAT_ERR_LAUNCH;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) return YOLAT_ERR_LAUNCH;
    *launches += 1;
    *total_ms += (double)ms;
  }
  return YOLAT_OK;
}

}  // extern "C"
