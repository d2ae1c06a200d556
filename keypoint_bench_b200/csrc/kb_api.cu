// Version / error-string entry points of the C ABI (include/kb_b200.h).
#include "kb_common.cuh"

extern "C" int kb_version(void) { return 100; }   // 0.1.0

extern "C" const char* kb_error_string(int code) {
    switch (code) {
        case KB_OK: return "ok";
        case KB_ERR_BAD_ARG: return "kb_b200: bad argument";
        case KB_ERR_WORKSPACE: return "kb_b200: workspace too small";
        case KB_ERR_UNSUPPORTED: return "kb_b200: unsupported size or mode";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "kb_b200: unknown error";
}

int kb_knobs[8] = {0, 1, 0, 0, 0, 0, 0, 0};

extern "C" int kb_debug_knob(int knob, int value) {
    if (knob < 1 || knob > 7) return KB_ERR_BAD_ARG;
    const int prev = kb_knobs[knob];
    kb_knobs[knob] = value;
    return prev;
}

int kb_sm_count(int* sms) {
    static int cached[64] = {0};
    int dev = 0;
    KB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || cached[dev] == 0) {
        int n = 0;
        KB_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) cached[dev] = n;
        *sms = n;
        return KB_OK;
    }
    *sms = cached[dev];
    return KB_OK;
}
