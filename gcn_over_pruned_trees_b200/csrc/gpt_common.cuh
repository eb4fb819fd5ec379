// Shared helpers for the gpt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#ifndef GPT_HOST_EMULATION
#include <map>
#include <mutex>
#include <utility>
#endif

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

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// A training step is a chain of ~15 short dependent kernels; between two of them the GPU otherwise idles for the
// launch latency of the next grid.  Kernels launched through gpt_launch() carry
// cudaLaunchAttributeProgrammaticStreamSerialization: the next grid may be set up and its CTAs made resident while the
// previous one is still running.  Every such kernel starts with GPT_PDL_ENTER(): `griddepcontrol.wait` blocks until
// the preceding grid has COMPLETED and its memory is visible (so nothing about data dependencies changes, reads and
// writes alike), `griddepcontrol.launch_dependents` lets the grid after this one start its own launch.  Captured
// into a CUDA graph these become programmatic edges.  GPT_PDL=0 in the environment launches without the attribute.
#ifdef GPT_HOST_EMULATION   // tests/emu: the kernels' source compiled for the host (no GPU in the build container)
#define GPT_PDL_ENTER() do { } while (0)
#else
#define GPT_PDL_ENTER()                                                   \
    do {                                                                  \
        asm volatile("griddepcontrol.wait;" ::: "memory");                \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   \
    } while (0)
#endif

// Kernels whose prologue touches nothing the preceding grid produces (shared-memory / tensor-memory set-up, reads of
// data that older grids wrote) trigger first, run the prologue, and only then wait -- the prologue overlaps the tail of
// the preceding grid.  Global WRITES always come after the wait (the preceding grid may still read a recycled buffer).
#ifdef GPT_HOST_EMULATION
#define GPT_PDL_TRIGGER() do { } while (0)
#define GPT_PDL_WAIT() do { } while (0)
#else
#define GPT_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define GPT_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#endif

extern int g_gpt_pdl;

template <typename... KArgs, typename... Args>
static inline cudaError_t gpt_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_gpt_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Opt `kernel` in to `smem` bytes of dynamic shared memory.  The attribute belongs to the (device, kernel) pair, so it
// is remembered per device: a second GPU driven from the same process gets its own call.
template <typename K>
static inline int gpt_smem_opt_in(K kernel, size_t smem) {
    if (smem <= 48 * 1024) return GPT_OK;
#ifndef GPT_HOST_EMULATION
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> have;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& h = have[std::make_pair(dev, reinterpret_cast<const void*>(kernel))];
    if (smem > h) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        h = smem;
    }
#endif
    return GPT_OK;
}

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
