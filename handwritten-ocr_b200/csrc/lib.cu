// Library-level entry points: version, error text, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace ocrb {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("OCRB_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
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
