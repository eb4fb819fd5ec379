// TEST INFRASTRUCTURE -- host versions of csrc/aggregate.cu's shared-memory / cp.async / packed-add accessors (the
// device versions are inline PTX).  A "32-bit shared-window address" is the low half of the host address; all "shared"
// objects of the host build live in one thread-local block, whose upper address bits smem_u32() records (a block that
// straddles a 4 GiB boundary would break this; the build's arrays are a few hundred KiB).  cp.async copies complete
// immediately.
#pragma once
#include <cstring>

namespace emu {
inline uintptr_t smem_hi = 0;     // upper 32 bits of the host addresses of "shared memory" (one thread-local block)
inline unsigned char* smem_ptr(uint32_t a) { return reinterpret_cast<unsigned char*>(smem_hi | (uintptr_t)a); }
}  // namespace emu

inline uint32_t smem_u32(const void* p) {
    const uintptr_t v = reinterpret_cast<uintptr_t>(p);
    emu::smem_hi = v & ~(uintptr_t)0xffffffffu;
    return (uint32_t)v;
}
inline void cp_async16(uint32_t smem_dst, const void* gsrc) { std::memcpy(emu::smem_ptr(smem_dst), gsrc, 16); }
inline void cp_async4(uint32_t smem_dst, const void* gsrc) { std::memcpy(emu::smem_ptr(smem_dst), gsrc, 4); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}

struct Pack4 {
    unsigned long long lo, hi;
};
inline void add_pk(Pack4& a, const Pack4 b) {
    float x[4], y[4];
    std::memcpy(x, &a, 16);
    std::memcpy(y, &b, 16);
    for (int i = 0; i < 4; ++i) x[i] += y[i];
    std::memcpy(&a, x, 16);
}
inline void mul_pk(Pack4& a, const float s) {
    float x[4];
    std::memcpy(x, &a, 16);
    for (int i = 0; i < 4; ++i) x[i] *= s;
    std::memcpy(&a, x, 16);
}
#define GPT_OPAQUE(ptr) ((void)(ptr))
inline float4 unpack(const Pack4 a) {
    float4 v;
    std::memcpy(&v, &a, 16);
    return v;
}
inline float4 lds128(uint32_t a) {
    float4 v;
    std::memcpy(&v, emu::smem_ptr(a), 16);
    return v;
}
inline Pack4 lds_pk(uint32_t a) {
    Pack4 v;
    std::memcpy(&v, emu::smem_ptr(a), 16);
    return v;
}
inline uint2 lds64(uint32_t a) {
    uint2 v;
    std::memcpy(&v, emu::smem_ptr(a), 8);
    return v;
}
inline uint32_t lds32(uint32_t a) {
    uint32_t v;
    std::memcpy(&v, emu::smem_ptr(a), 4);
    return v;
}
inline uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    std::memcpy(&v, emu::smem_ptr(a), 2);
    return v;
}
inline void sts_u8(uint32_t a, uint32_t v) { *emu::smem_ptr(a) = (unsigned char)v; }
inline void sts128(uint32_t a, const float4 v) { std::memcpy(emu::smem_ptr(a), &v, 16); }
inline void sts_f32(uint32_t a, float v) { std::memcpy(emu::smem_ptr(a), &v, 4); }
inline void stg128(void* gptr, const float4 v) { std::memcpy(gptr, &v, 16); }
inline void fence_mbar_init() {}
