// Library-level entry points of libocrpp.so: version, error reporting, launch accounting.
#include "common.cuh"

#include <cstdlib>

namespace ocrpp {

std::atomic<long long> g_launch_count{0};

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

namespace {
std::atomic<int> g_tuning[OCRPP_TUNE_COUNT];
}  // namespace
int tuning(int key) { return (key >= 0 && key < OCRPP_TUNE_COUNT) ? g_tuning[key].load(std::memory_order_relaxed) : 0; }

bool debug_sync() {
  static const bool on = [] {
    const char* e = getenv("OCRPP_DEBUG_SYNC");
    return e && e[0] == '1';
  }();
  return on;
}

// ---- profiling ring --------------------------------------------------------------------------
namespace {
bool g_prof_on = false;
int g_prof_calls = 0;
int g_prof_phases = 0;
const char* g_prof_names[kProfMaxPhases] = {nullptr};
cudaEvent_t g_prof_ev[kProfMaxCalls][kProfMaxPhases + 1];
bool g_prof_ev_ready = false;
}  // namespace

bool profile_on() { return g_prof_on; }

int profile_begin(cudaStream_t s) {
  if (!g_prof_on || g_prof_calls >= kProfMaxCalls) return -1;
  if (!g_prof_ev_ready) {
    for (int c = 0; c < kProfMaxCalls; ++c)
      for (int p = 0; p <= kProfMaxPhases; ++p) cudaEventCreate(&g_prof_ev[c][p]);
    g_prof_ev_ready = true;
  }
  const int slot = g_prof_calls++;
  cudaEventRecord(g_prof_ev[slot][0], s);
  return slot;
}

void profile_mark(int slot, int phase, const char* name, cudaStream_t s) {
  if (slot < 0 || phase >= kProfMaxPhases) return;
  cudaEventRecord(g_prof_ev[slot][phase + 1], s);
  g_prof_names[phase] = name;
  if (phase + 1 > g_prof_phases) g_prof_phases = phase + 1;
}

}  // namespace ocrpp

extern "C" {

void ocrpp_profile_enable(int on) { ocrpp::g_prof_on = on != 0; }
void ocrpp_profile_reset(void) {
  ocrpp::g_prof_calls = 0;
  ocrpp::g_prof_phases = 0;   // the next profiled call defines the phase list (another entry point may follow)
}
int ocrpp_profile_read(float* ms_out, int cap, int* calls_out) {
  using namespace ocrpp;
  const int np = g_prof_phases < cap ? g_prof_phases : cap;
  for (int p = 0; p < np; ++p) {
    float sum = 0.f;
    for (int c = 0; c < g_prof_calls; ++c) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, g_prof_ev[c][p], g_prof_ev[c][p + 1]) == cudaSuccess) sum += ms;
    }
    ms_out[p] = sum;
  }
  if (calls_out) *calls_out = g_prof_calls;
  return np;
}
const char* ocrpp_profile_phase_name(int phase) {
  using namespace ocrpp;
  return (phase >= 0 && phase < g_prof_phases && g_prof_names[phase]) ? g_prof_names[phase] : "";
}


int ocrpp_set_tuning(int key, int value) {
  using namespace ocrpp;
  OCRPP_CHECK_ARG(key >= 0 && key < OCRPP_TUNE_COUNT, "set_tuning: unknown key %d", key);
  g_tuning[key].store(value);
  return OCRPP_OK;
}

int ocrpp_abi_version(void) { return OCRPP_ABI_VERSION; }
const char* ocrpp_last_error(void) { return ocrpp::last_error_buf(); }
int64_t ocrpp_launch_count(void) { return (int64_t)ocrpp::g_launch_count.load(); }
void ocrpp_reset_launch_count(void) { ocrpp::g_launch_count.store(0); }

}  // extern "C"
