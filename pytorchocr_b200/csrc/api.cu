// Library-level entry points of libocrpp.so: version, error reporting, launch accounting.
#include "common.cuh"

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

}  // namespace ocrpp

extern "C" {

int ocrpp_abi_version(void) { return OCRPP_ABI_VERSION; }
const char* ocrpp_last_error(void) { return ocrpp::last_error_buf(); }
int64_t ocrpp_launch_count(void) { return (int64_t)ocrpp::g_launch_count.load(); }
void ocrpp_reset_launch_count(void) { ocrpp::g_launch_count.store(0); }

}  // extern "C"
