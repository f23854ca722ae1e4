// nnk_core.cu -- error text, launch counter, ABI version of libnnk_b200.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "nnk_common.cuh"

namespace nnk {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace nnk

extern "C" const char* nnk_last_error(void) { return nnk::g_err; }
extern "C" int nnk_abi_version(void) { return NNK_ABI_VERSION; }
extern "C" int64_t nnk_launch_count(void) { return (int64_t)nnk::g_launches.load(); }
extern "C" void nnk_status_decode(uint64_t word, nnk_status_t* out) {
  if (out) nnk::decode_status((unsigned long long)word, out);
}
