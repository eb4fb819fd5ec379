// Library-level entry points of libgptb200.so (include/gpt_b200.h).
#include "gpt_common.cuh"

unsigned long long g_gpt_launches = 0;

#include <cstdlib>
int g_gpt_pdl = [] {
    const char* e = getenv("GPT_PDL");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
}();

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

// Pull a buffer into L2 ahead of its first use (one prefetch per 128-byte line, no data returned to the SM).
// engine.FusedTrainStep issues this for the flat parameter buffer on a side stream at the start of a step, so the
// classifier head and the projections do not pay DRAM latency on their first touch of the weights.
__global__ void l2_prefetch_kernel(const unsigned char* __restrict__ p, size_t lines) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < lines; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128));
}

extern "C" int gpt_l2_prefetch(const void* ptr, long long bytes, void* stream) {
    GPT_CHECK_ARG(ptr != nullptr && bytes >= 0);
    if (bytes == 0) return GPT_OK;
    const size_t lines = ((size_t)bytes + 127) / 128;
    size_t blocks = (lines + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    l2_prefetch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const unsigned char*>(ptr), lines);
    return gpt_launch_status();
}
