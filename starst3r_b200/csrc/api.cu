// Error reporting and device queries shared by every C-ABI entry point.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "../../include/starst3r_b200.h"

static thread_local char g_err[512] = "";
unsigned long long g_st3r_launches = 0;

void st3r_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int st3r_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" {
const char* st3r_last_error(void) { return g_err; }
int st3r_abi_version(void) { return ST3R_ABI_VERSION; }
int st3r_device_sm_count(void) { return st3r_num_sms(); }
uint64_t st3r_launch_count(void) { return g_st3r_launches; }
void st3r_launch_count_add(uint64_t n) { g_st3r_launches += n; }
}
