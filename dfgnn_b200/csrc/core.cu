// core.cu -- error reporting / bookkeeping for the C ABI (include/dfgnn_b200.h).
#include <stdarg.h>
#include <string.h>

#include "abi_common.h"

namespace dfgnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// process-wide (autograd runs the backward on its own thread, the caller asks from another)
static std::atomic<const char*> g_last_kernel[3] = {{""}, {""}, {""}};

void note_kernel(int slot, const char* name) {
  if (slot >= 0 && slot < 3) g_last_kernel[slot].store(name, std::memory_order_relaxed);
}

std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

}  // namespace dfgnn

extern "C" {

int dfgnn_abi_version(void) { return DFGNN_ABI_VERSION; }
const char* dfgnn_last_error(void) { return dfgnn::g_err; }
uint64_t dfgnn_launch_count(void) { return dfgnn::launch_counter().load(); }
const char* dfgnn_last_kernel(int slot) {
  return (slot >= 0 && slot < 3) ? dfgnn::g_last_kernel[slot].load(std::memory_order_relaxed) : "";
}

}  // extern "C"
