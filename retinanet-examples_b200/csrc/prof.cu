// prof.cu -- see prof.cuh.  Event pairs are kept per tag and summed on odtk_prof_get().
#include <vector>

#include "common.cuh"
#include "prof.cuh"

namespace {
struct Pair { cudaEvent_t a, b; };
bool g_on = false;
std::vector<Pair> g_pairs[ODTK_PROF_NTAGS];
std::vector<Pair> g_pool;
cudaEvent_t g_open[ODTK_PROF_NTAGS];
bool g_is_open[ODTK_PROF_NTAGS] = {false};
constexpr size_t kMaxPairs = 1 << 16;
}  // namespace

void odtk_prof_begin(int tag, cudaStream_t s) {
  if (!g_on || tag < 0 || tag >= ODTK_PROF_NTAGS || g_pairs[tag].size() >= kMaxPairs) return;
  Pair p;
  if (!g_pool.empty()) { p = g_pool.back(); g_pool.pop_back(); }
  else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
  cudaEventRecord(p.a, s);
  g_pairs[tag].push_back(p);
  g_is_open[tag] = true;
}

void odtk_prof_end(int tag, cudaStream_t s) {
  if (!g_on || tag < 0 || tag >= ODTK_PROF_NTAGS || !g_is_open[tag]) return;
  cudaEventRecord(g_pairs[tag].back().b, s);
  g_is_open[tag] = false;
}

extern "C" void odtk_prof_enable(int on) { g_on = on != 0; }

extern "C" void odtk_prof_reset(void) {
  for (int t = 0; t < ODTK_PROF_NTAGS; t++) {
    for (auto &p : g_pairs[t]) g_pool.push_back(p);
    g_pairs[t].clear();
    g_is_open[t] = false;
  }
}

// Synchronises the device, then returns the summed duration (ms) and the number of timed launches.
extern "C" int odtk_prof_get(int tag, double *total_ms, long long *launches) {
  if (tag < 0 || tag >= ODTK_PROF_NTAGS || !total_ms || !launches) return ODTK_E_INVALID;
  if (cudaDeviceSynchronize() != cudaSuccess) return ODTK_E_CUDA;
  double tot = 0;
  long long n = 0;
  for (auto &p : g_pairs[tag]) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { tot += ms; n++; }
  }
  *total_ms = tot;
  *launches = n;
  return ODTK_OK;
}

// Per-launch durations (ms) of a tag in launch order (the last `cap` launches when there are more): bench.py joins the
// conv list of one eager step with the engine trace to show where the time goes INSIDE a running step (sustained clocks),
// which the cold, serialised ncu launch list cannot.
extern "C" long long odtk_prof_get_list(int tag, float *ms, long long cap) {
  if (tag < 0 || tag >= ODTK_PROF_NTAGS || !ms || cap <= 0) return ODTK_E_INVALID;
  if (cudaDeviceSynchronize() != cudaSuccess) return ODTK_E_CUDA;
  const long long n = (long long)g_pairs[tag].size();
  const long long first = n > cap ? n - cap : 0;
  long long k = 0;
  for (long long i = first; i < n; i++) {
    float v = 0;
    if (cudaEventElapsedTime(&v, g_pairs[tag][i].a, g_pairs[tag][i].b) != cudaSuccess) v = -1.0f;
    ms[k++] = v;
  }
  return k;
}
