// TEST INFRASTRUCTURE -- a host stand-in for <cuda_runtime.h>, just large enough to compile csrc/deprel.cu,
// csrc/prune_csr.cu, csrc/pool3.cu, csrc/gemm_simt.cu, csrc/embed.cu, csrc/aggregate.cu, csrc/update.cu and csrc/batch.cu (and the helpers of csrc/gpt_common.cuh they use) with g++ and run its kernels on the CPU: every CUDA thread of a block is a
// fiber (ucontext) on the calling OS thread, scheduled round-robin; __syncthreads() and the warp shuffles are barriers
// at which a fiber yields; blocks run one after the other; exited threads stop counting towards barriers, as on the
// device.  Single-threaded and deterministic (atomics are plain adds).
// The build container has no GPU; this lets `-m "not gpu"` tests execute the kernels' actual source against the oracle.
// Nothing under gcn_over_pruned_trees_b200/ includes or links this.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <utility>
#include <vector>

inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }

#define __global__
#define __grid_constant__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ thread_local   // one OS thread runs every fiber: function-scope and extern shared arrays alike
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct emu_uint3 {
    unsigned x, y, z;
};
struct __attribute__((aligned(16))) float4 {
    float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct __attribute__((aligned(8))) uint2 {
    unsigned x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline unsigned __float_as_uint(float v) {
    unsigned u;
    std::memcpy(&u, &v, 4);
    return u;
}
inline float __uint_as_float(unsigned u) {
    float v;
    std::memcpy(&v, &u, 4);
    return v;
}
inline void __trap() { std::abort(); }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
constexpr unsigned long long cudaEnableDefault = 0;
inline int cudaGetDriverEntryPoint(const char*, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
    *fn = nullptr;
    *q = cudaDriverEntryPointSymbolNotFound;     // no driver on the host: no tensor maps, K2 takes its cp.async path
    return 1;
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline float __frcp_rn(float x) { return 1.0f / x; }

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

inline emu_uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emu {
constexpr size_t kStackBytes = 128 * 1024;

struct Barrier {
    int arrived = 0, live = 0;
    unsigned generation = 0;
};
#if defined(__x86_64__)
#define EMU_FAST_SWITCH 1      // hand-written stack switch (emu_switch.cpp): swapcontext() costs two system calls
extern "C" void emu_switch(void** save_sp, void* const* load_sp);
#endif
struct Fiber {
    ucontext_t ctx;
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
};
struct Block {
    int n = 0, current = 0;
    std::vector<Fiber> fibers;
    ucontext_t scheduler;
    void* scheduler_sp = nullptr;
    Barrier sync;
    std::vector<Barrier> warp;
    std::vector<uint32_t> warp_buf;
    std::function<void()> body;
};
inline Block* block = nullptr;

inline void yield() {
#ifdef EMU_FAST_SWITCH
    emu_switch(&block->fibers[block->current].sp, &block->scheduler_sp);
#else
    swapcontext(&block->fibers[block->current].ctx, &block->scheduler);
#endif
}

inline void arrive_and_wait(Barrier& b) {
    const unsigned g = b.generation;
    if (++b.arrived >= b.live) {
        b.arrived = 0;
        ++b.generation;
        return;
    }
    while (b.generation == g) yield();
}
// a thread that has exited no longer takes part in barriers
inline void leave(Barrier& b) {
    --b.live;
    if (b.live > 0 && b.arrived >= b.live) {
        b.arrived = 0;
        ++b.generation;
    }
}
inline void trampoline() {
    Block* blk = block;
    blk->body();
    blk->fibers[blk->current].done = true;
    leave(blk->sync);
    leave(blk->warp[blk->current >> 5]);
#ifdef EMU_FAST_SWITCH
    emu_switch(&blk->fibers[blk->current].sp, &blk->scheduler_sp);     // never resumed
    std::abort();
#endif
}

inline void run_block(Block& blk) {
    const int n = blk.n;
    blk.sync = Barrier{0, n, 0};
    for (int w = 0; w < (int)blk.warp.size(); ++w) blk.warp[w] = Barrier{0, (w + 1) * 32 <= n ? 32 : n - w * 32, 0};
    for (int t = 0; t < n; ++t) {
        Fiber& f = blk.fibers[t];
        f.done = false;
#ifdef EMU_FAST_SWITCH
        // frame emu_switch() pops: six callee-saved registers, then `ret` into trampoline with the stack as after a call
        void** top = reinterpret_cast<void**>((reinterpret_cast<uintptr_t>(f.stack) + kStackBytes) & ~(uintptr_t)15);
        top[-1] = nullptr;
        top[-2] = reinterpret_cast<void*>(&trampoline);
        for (int i = 3; i <= 8; ++i) top[-i] = nullptr;
        f.sp = top - 8;
#else
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = &blk.scheduler;
        makecontext(&f.ctx, trampoline, 0);
#endif
    }
    for (int remaining = n; remaining > 0;) {
        remaining = 0;
        for (int t = 0; t < n; ++t) {
            Fiber& f = blk.fibers[t];
            if (f.done) continue;
            blk.current = t;
            threadIdx = emu_uint3{(unsigned)t, 0, 0};
#ifdef EMU_FAST_SWITCH
            emu_switch(&blk.scheduler_sp, &f.sp);
#else
            swapcontext(&blk.scheduler, &f.ctx);
#endif
            if (!f.done) ++remaining;
        }
    }
}
}  // namespace emu

inline void __syncthreads() {
    emu::arrive_and_wait(emu::block->sync);
    threadIdx = emu_uint3{(unsigned)emu::block->current, 0, 0};
}

template <typename T>
inline T emu_shfl(T v, int src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const int me = emu::block->current, warp = me >> 5, lane = me & 31;
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    emu::block->warp_buf[(size_t)warp * 32 + lane] = bits;
    emu::arrive_and_wait(emu::block->warp[warp]);
    const uint32_t got = emu::block->warp_buf[(size_t)warp * 32 + (src_lane & 31)];
    emu::arrive_and_wait(emu::block->warp[warp]);
    threadIdx = emu_uint3{(unsigned)me, 0, 0};
    T r;
    std::memcpy(&r, &got, 4);
    return r;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_shfl(v, (emu::block->current & 31) ^ o); }
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int o) {
    const int lane = emu::block->current & 31;
    return emu_shfl(v, lane >= o ? lane - o : lane);
}

inline void __syncwarp(unsigned = 0xffffffffu) {
    const int me = emu::block->current;
    emu::arrive_and_wait(emu::block->warp[me >> 5]);
    threadIdx = emu_uint3{(unsigned)me, 0, 0};
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) { return emu_shfl(v, src_lane); }
// lanes of the warp that have not exited (K1's warps leave together, so this is 32 or the tail of a short block)
inline int emu_warp_lanes() {
    const int warp = emu::block->current >> 5;
    return (warp + 1) * 32 <= emu::block->n ? 32 : emu::block->n - warp * 32;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    const int me = emu::block->current, warp = me >> 5, lane = me & 31;
    emu::block->warp_buf[(size_t)warp * 32 + lane] = pred ? 1u : 0u;
    emu::arrive_and_wait(emu::block->warp[warp]);
    unsigned m = 0;
    for (int l = 0; l < emu_warp_lanes(); ++l) m |= emu::block->warp_buf[(size_t)warp * 32 + l] << l;
    emu::arrive_and_wait(emu::block->warp[warp]);
    threadIdx = emu_uint3{(unsigned)me, 0, 0};
    return m;
}
inline unsigned __match_any_sync(unsigned, int v) {
    const int me = emu::block->current, warp = me >> 5, lane = me & 31;
    emu::block->warp_buf[(size_t)warp * 32 + lane] = (uint32_t)v;
    emu::arrive_and_wait(emu::block->warp[warp]);
    unsigned m = 0;
    for (int l = 0; l < emu_warp_lanes(); ++l) m |= (emu::block->warp_buf[(size_t)warp * 32 + l] == (uint32_t)v ? 1u : 0u) << l;
    emu::arrive_and_wait(emu::block->warp[warp]);
    threadIdx = emu_uint3{(unsigned)me, 0, 0};
    return m;
}
inline int __reduce_max_sync(unsigned, int v) {
    const int me = emu::block->current, warp = me >> 5, lane = me & 31;
    emu::block->warp_buf[(size_t)warp * 32 + lane] = (uint32_t)v;
    emu::arrive_and_wait(emu::block->warp[warp]);
    int m = v;
    for (int l = 0; l < emu_warp_lanes(); ++l) m = max(m, (int)emu::block->warp_buf[(size_t)warp * 32 + l]);
    emu::arrive_and_wait(emu::block->warp[warp]);
    threadIdx = emu_uint3{(unsigned)me, 0, 0};
    return m;
}
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int atomicMin(int* p, int v) {
    const int old = *p;
    if (v < old) *p = v;
    return old;
}
inline int atomicAdd(int* p, int v) {
    const int old = *p;
    *p = old + v;
    return old;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

inline float atomicAdd(float* p, float v) {
    const float old = *p;
    *p = old + v;
    return old;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

template <typename... K, typename... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(K...), A... args) {
    const int n = (int)cfg->blockDim.x;
    if (cfg->blockDim.y != 1 || cfg->blockDim.z != 1 || n < 1) return 1;
    emu::Block blk;
    blk.n = n;
    blk.fibers.resize(n);
    blk.warp.resize((n + 31) / 32);
    blk.warp_buf.resize((size_t)((n + 31) / 32) * 32);
    for (auto& f : blk.fibers) f.stack = static_cast<char*>(std::malloc(emu::kStackBytes));
    blk.body = [&]() { kernel(args...); };
    emu::block = &blk;
    blockDim = cfg->blockDim;
    gridDim = cfg->gridDim;
    for (unsigned bz = 0; bz < cfg->gridDim.z; ++bz)
        for (unsigned by = 0; by < cfg->gridDim.y; ++by)
            for (unsigned bx = 0; bx < cfg->gridDim.x; ++bx) {
                blockIdx = emu_uint3{bx, by, bz};
                emu::run_block(blk);      // shared memory belongs to one block at a time
            }
    for (auto& f : blk.fibers) std::free(f.stack);
    emu::block = nullptr;
    return cudaSuccess;
}
