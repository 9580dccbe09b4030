// Library-level pieces of the C-ABI: error string, version, launch counter.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static char g_err[512] = "";
extern "C" unsigned long long npb_launch_counter = 0;
extern "C" double npb_bytes_counter = 0.0;

void npb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

const char* npb_last_error(void) { return g_err; }

int npb_version(void) { return 100; }

// number of kernels launched through this library since load / last reset
unsigned long long npb_launch_count(void) { return npb_launch_counter; }
void npb_launch_count_reset(void) { npb_launch_counter = 0; }
double npb_bytes_count(void) { return npb_bytes_counter; }
void npb_bytes_count_reset(void) { npb_bytes_counter = 0.0; }

// 0 when a CUDA device is usable; the Python layer refuses to run without it
int npb_device_check(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        npb_set_error("no CUDA device: %s", cudaGetErrorString(e));
        return NPB_ERR_CUDA;
    }
    return NPB_OK;
}

}  // extern "C"
