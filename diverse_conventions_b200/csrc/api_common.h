// api_common.h — helpers shared by the C-ABI translation units
#pragma once
#include <cuda_runtime.h>

namespace ocb {

extern thread_local char g_err[512];
// records a printf-style message for ocb_last_error() and returns `code`
int fail(int code, const char* fmt, ...);

// makes `device` current for the scope of one ABI call and restores the caller's device
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) {
            cudaSetDevice(device);
            switched = true;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

// A handle may be destroyed (e.g. by a garbage collector) while ANOTHER stream of this thread is being
// captured into a CUDA graph; cudaFree is a "potentially unsafe" call that would invalidate a
// global-mode capture.  Relaxed mode for the scope of the frees keeps the capture intact.
struct CaptureRelaxed {
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    CaptureRelaxed() { cudaThreadExchangeStreamCaptureMode(&mode); }
    ~CaptureRelaxed() { cudaThreadExchangeStreamCaptureMode(&mode); }
};

}  // namespace ocb
