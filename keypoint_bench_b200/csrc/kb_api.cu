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
