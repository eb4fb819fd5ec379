// TEST INFRASTRUCTURE -- a host stand-in for <cuda_runtime.h>, just large enough to compile csrc/deprel.cu (and the
// helpers of csrc/gpt_common.cuh it uses) with g++ and run its kernels on CPU threads: one std::thread per CUDA thread of
// a block, blocks executed one after the other, __syncthreads() / warp shuffles as barriers, atomics under a mutex.
// The build container has no GPU; this lets `-m "not gpu"` tests execute the kernels' actual source against the oracle.
// Nothing under gcn_over_pruned_trees_b200/ includes or links this.
#pragma once
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct emu_uint3 {
    unsigned x, y, z;
};

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef void* cudaStream_t;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }

constexpr int cudaLaunchAttributeProgrammaticStreamSerialization = 1;
struct cudaLaunchAttribute {
    int id;
    struct {
        int programmaticStreamSerializationAllowed;
    } val;
};
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute* attrs;
    unsigned numAttrs;
};

namespace emu {
struct Block {
    std::barrier<> sync_bar, block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<uint32_t> warp_buf;
    explicit Block(int n) : sync_bar(n), block_bar(n), warp_buf((size_t)((n + 31) / 32) * 32) {
        for (int w = 0; w < (n + 31) / 32; ++w) {
            const int lanes = (w + 1) * 32 <= n ? 32 : n - w * 32;
            warp_bar.emplace_back(new std::barrier<>(lanes));
        }
    }
};
inline thread_local Block* block = nullptr;
inline std::mutex atomic_mutex;
}  // namespace emu

inline thread_local emu_uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { emu::block->sync_bar.arrive_and_wait(); }

template <typename T>
inline T emu_shfl(T v, int src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    emu::block->warp_buf[(size_t)warp * 32 + lane] = bits;
    emu::block->warp_bar[warp]->arrive_and_wait();
    const uint32_t got = emu::block->warp_buf[(size_t)warp * 32 + (src_lane & 31)];
    emu::block->warp_bar[warp]->arrive_and_wait();
    T r;
    std::memcpy(&r, &got, 4);
    return r;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_shfl(v, (int)(threadIdx.x & 31) ^ o); }
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int o) {
    const int lane = threadIdx.x & 31;
    return emu_shfl(v, lane >= o ? lane - o : lane);
}

inline float atomicAdd(float* p, float v) {
    std::lock_guard<std::mutex> g(emu::atomic_mutex);
    const float old = *p;
    *p = old + v;
    return old;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }

template <typename... K, typename... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(K...), A... args) {
    const int n = (int)cfg->blockDim.x;
    if (cfg->blockDim.y != 1 || cfg->blockDim.z != 1 || cfg->gridDim.z != 1 || n < 1) return 1;
    emu::Block blk(n);
    const dim3 grid = cfg->gridDim, bdim = cfg->blockDim;
    std::vector<std::thread> threads;
    for (int t = 0; t < n; ++t) {
        threads.emplace_back([&, t]() {
            emu::block = &blk;
            threadIdx = emu_uint3{(unsigned)t, 0, 0};
            blockDim = bdim;
            gridDim = grid;
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    blockIdx = emu_uint3{bx, by, 0};
                    kernel(args...);
                    blk.block_bar.arrive_and_wait();     // shared memory belongs to one block at a time
                }
        });
    }
    for (auto& th : threads) th.join();
    return cudaSuccess;
}
