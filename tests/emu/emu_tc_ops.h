// TEST INFRASTRUCTURE -- stand-ins for the mbarrier / TMA entry points of csrc/tcgen05_util.cuh in the host build.
// The host build never creates a tensor map (cudaGetDriverEntryPoint fails in tests/emu/cuda_runtime.h), so K2 takes
// its cp.async path and none of these is reached; they only have to compile.
#pragma once
#include <cstdlib>

inline uint32_t smem_addr(const void* p) { return (uint32_t)(uintptr_t)p; }
inline void mbar_init(uint32_t, uint32_t) {}
inline void mbar_expect_tx(uint32_t, uint32_t) { std::abort(); }
inline void mbar_wait(uint32_t, uint32_t) { std::abort(); }
inline void mbar_arrive(uint32_t) { std::abort(); }
inline void tma_load_2d(uint32_t, const CUtensorMap*, uint32_t, int, int) { std::abort(); }
inline void prefetch_map(const CUtensorMap*) {}
inline void fence_async_smem() {}
