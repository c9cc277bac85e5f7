// pcb200 — error plumbing, version and device probe for the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/pcb200.h"
#include "pcb_common.cuh"

namespace pcb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED); }
}  // namespace pcb

extern "C" const char* pcb_last_error(void) { return pcb::g_err; }
extern "C" int pcb_version(void) { return 100; }
extern "C" int64_t pcb_launch_count(void) { return (int64_t)__atomic_load_n(&pcb::g_launches, __ATOMIC_RELAXED); }
extern "C" int pcb_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { pcb::set_error("no CUDA device"); cudaGetLastError(); return 0; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { pcb::set_error("cudaGetDeviceProperties failed"); cudaGetLastError(); return 0; }
  if (p.major != 10) { pcb::set_error("pcb200 kernels are sm_100a only; device is cc %d.%d", p.major, p.minor); return 0; }
  return 1;
}
