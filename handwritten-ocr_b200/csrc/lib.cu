// Library-level entry points: version, error text, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace ocrb {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
bool pdl_enabled(int bit) {
  static int mask = -1;
  if (mask < 0) {
    const char *e = getenv("OCRB_PDL");      // bit mask: 1 = skinny GEMM, 2 = decode attention, 4 = attention combine
    mask = e ? atoi(e) : 7;
  }
  return (mask & bit) != 0;
}
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ocrb

extern "C" {
int ocrb_version(void) { return 1; }
const char *ocrb_last_error(void) { return ocrb::g_err; }
uint64_t ocrb_launch_count(void) { return ocrb::g_launches.load(); }
void ocrb_launch_count_reset(void) { ocrb::g_launches.store(0); }
}
