// Library-level entry points of libgptb200.so (include/gpt_b200.h).
#include "gpt_common.cuh"

unsigned long long g_gpt_launches = 0;

extern "C" int gpt_version(void) { return 100; }

extern "C" unsigned long long gpt_launch_count(void) { return g_gpt_launches; }

extern "C" const char* gpt_error_string(int code) {
    switch (code) {
        case GPT_OK: return "ok";
        case GPT_ERR_BAD_ARG: return "gpt_b200: bad argument (null pointer or size out of range)";
        case GPT_ERR_UNSUPPORTED: return "gpt_b200: shape not supported by this kernel";
        case GPT_ERR_DRIVER: return "gpt_b200: CUDA driver entry point unavailable";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "gpt_b200: unknown error";
}
