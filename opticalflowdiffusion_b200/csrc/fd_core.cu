// libflowdiff.so: library info, error reporting, device check.
#include "fd_common.cuh"

#include <string.h>

static thread_local char g_fd_err[512] = "";
unsigned long long g_fd_launches = 0;

void fd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_fd_err, sizeof(g_fd_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int fd_version(void) { return 100; }

const char* fd_arch(void) { return "sm_100a"; }

const char* fd_last_error(void) { return g_fd_err; }

int fd_device_check(void) {
  int dev = 0;
  FD_CUDA(cudaGetDevice(&dev));
  int major = 0;
  FD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    fd_set_error("libflowdiff is built for sm_100a only; current device has compute capability %d.x", major);
    return FD_EARCH;
  }
  return FD_OK;
}

unsigned long long fd_launch_count(void) { return g_fd_launches; }

int fd_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

}  // extern "C"
