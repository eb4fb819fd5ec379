// Shared helpers for the gpt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GPT_FULL_MASK 0xffffffffu

// Return codes of every extern "C" entry point (include/gpt_b200.h).
#define GPT_OK 0
#define GPT_ERR_BAD_ARG (-1)
#define GPT_ERR_UNSUPPORTED (-2)
#define GPT_ERR_DRIVER (-3)
// positive values are cudaError_t

#define GPT_CHECK_ARG(cond) \
    do {                    \
        if (!(cond)) return GPT_ERR_BAD_ARG; \
    } while (0)

// number of kernels this library has launched (host-side counter, read through gpt_launch_count())
extern unsigned long long g_gpt_launches;

// call once after every <<<>>> launch
static inline int gpt_launch_status() {
    ++g_gpt_launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GPT_OK : (int)e;
}

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GPT_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(GPT_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ int warp_or_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(GPT_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GPT_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(GPT_FULL_MASK, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// Philox4x32-10 (Salmon et al. 2011): counter-based RNG used for in-kernel dropout.
struct Philox4 {
    uint32_t x, y, z, w;
};
// kPhiloxRounds = 7 is the smallest round count that is Crush-resistant in the paper (10 is its safety default);
// dropout masks do not need the margin and the rounds are the kernel's largest ALU cost.
constexpr int kPhiloxRounds = 7;
__device__ __forceinline__ Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}
// uniform in [0,1) from 32 random bits (24-bit mantissa)
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// flags[b,t] bits written by gpt_prune_csr
#define GPT_FLAG_INTREE 1u  // token has >= 1 adjacency entry: NOT pool-masked (model/gcn.py:262)
#define GPT_FLAG_SUBJ 2u    // subj_pos == 0 (model/gcn.py:116)
#define GPT_FLAG_OBJ 4u     // obj_pos == 0
