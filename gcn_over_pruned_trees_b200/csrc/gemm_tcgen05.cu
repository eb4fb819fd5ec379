// K3 (tensor-core mode) -- the W projection of a GCN layer as a tcgen05 / TMEM GEMM fed by TMA (sm_100a).
//
//   Y[M,N] = X[M,K] . W[N,K]^T     fp32 in HBM, TF32 on the 5th-gen tensor cores, fp32 accumulation in TMEM
//
// (the one dense contraction on the path: /root/reference/model/gcn.py:270-271 computes W(Ax) + W(h); by linearity a
// single projection per layer suffices, see aggregate.cu).  Both operands are K-major, so fp32 rows are loaded as
// they lie in HBM: TMA brings [128 x 32] / [N_tile x 32] fp32 boxes (128-byte rows, SWIZZLE_128B) into a 4-stage
// shared-memory ring, one elected thread issues tcgen05.mma.kind::tf32 (M = 128, N = N_tile <= 256, K = 8 per
// instruction, 4 per stage) with the accumulator in tensor memory, tcgen05.commit releases ring slots and finally
// signals the epilogue warps, which read the accumulator with tcgen05.ld (32 lanes x 32 columns per warp) and store
// fp32 rows.  No operand conversion pass: kind::tf32 consumes the fp32 bit patterns (low 13 mantissa bits ignored).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (warp w may only touch TMEM lanes 32*(w % 4) .. +31, so the four epilogue warps cover all 128 rows).
//
// dgrad  dX[M,K] = dY[M,N] . W[N,K]  runs through the same kernel with a transposed copy of W (tiny) as operand B.
//
// Accuracy modes.  kind::tf32 truncates each fp32 operand to its upper 19 bits (measured on B200: the low 13 mantissa
// bits are ignored, for A and for B), so with  a = a_hi + a_lo,  a_hi = trunc_tf32(a),  a_lo = a - a_hi (exact):
//   PASSES = 1:  A.B ~ A_hi.B_hi                                   one TF32 pass, ~1e-3 relative
//   PASSES = 3:  A.B ~ A_hi.B_hi + A_lo.B_hi + A_hi.B_lo           "3xTF32": fp32-grade (~1e-6), the parity mode
//   PASSES = 0:  A.B ~ bf16(A).bf16(B), fp32 accumulation           kind::f16 with bf16 operands (the reduced-precision
//                mode north_star asks for; ~4e-3 relative per product): the landed fp32 X tile is rounded to bf16 in
//                shared memory by the same warps (in place: [128 x 32] bf16, 64-byte rows, SWIZZLE_64B), W arrives as bf16
// For PASSES = 3 hi is *rounded* to TF32 (|lo| <= 2^-12 |a|, zero-mean), which makes the truncation the tensor core
// applies to lo and the dropped lo.lo term ~2^-23 each.  B_hi / B_lo (the weight) are precomputed by a tiny kernel
// and arrive by TMA; A_hi / A_lo are produced in shared memory by the four epilogue warps, which are idle during the
// main loop: they rewrite the freshly landed A tile in place as hi and store lo next to it (element-wise, so the
// 128B swizzle is preserved), fence the generic->async proxy and arrive on a per-stage mbarrier the MMA thread
// waits on.
#include "gpt_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int BM = 128;        // rows per CTA tile (UMMA M)
constexpr int BK = 32;         // fp32 elements per k-block = one 128-byte swizzle atom
constexpr int UMMA_K = 8;      // tf32: 32 bytes per instruction
constexpr int kMaxStages = 8;   // ring depth is chosen at launch: as many stages as fit in shared memory
constexpr int kMaxChainKb = 32; // split-K: k-blocks per accumulator chain (bounds the round-toward-zero bias, run_tf32_gemm)
constexpr int kSplitWarps = 8;          // warps 2..9: operand split (3xTF32) during the main loop, then the epilogue
constexpr int kGemmThreads = 64 + 32 * kSplitWarps;

// Optional epilogue of the data-gradient GEMM: the consumer of dX is K2's backward of the previous layer, whose first
// step is g = dX * dropscale * [out > 0] / denom.  Doing it here, where every thread already holds 32 consecutive
// columns of one row, removes a whole shared-memory pass (and a barrier per slice) from the HBM-bound K2 backward.
struct MaskEpilogue {
    const uint32_t* act;     // activation bits of the previous layer's forward, [B, ceil(N/32), T] (K2 layout); or null
    const float* denom;      // [B*T]
    float scale;             // dropout scale of the previous layer's forward
    int T;
    // Unrelated to the mask, carried here because every launcher already passes this struct: an optional device-side row
    // count (gpt_live_rows).  Row tiles at or beyond *m_live leave at once -- the relation-aware layers project only the
    // compacted observable rows of a batch, and inside a captured step the host never knows how many there are.
    const int* m_live;
    // optional column bias added to the accumulator before it is stored (y = x W^T + b; never with split-K): the relation
    // mix reads every projected element once per direction, and a bias left for it to add doubles its L2 traffic
    const float* bias;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();  // a lost arrival must fail the launch, never hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) {  // round to nearest TF32 (ties away), low 13 bits zero
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major, canonical value 1) in [16,30), stride byte offset =
// 8 rows x 128 B = 1024 B (>> 4 = 64) in [32,46), descriptor version 1 (Blackwell) in [46,48), layout 2 = SWIZZLE_128B
// in [61,64).  The tile base must be 1024-byte aligned; advancing K inside the atom adds (bytes >> 4) to the address.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major SWIZZLE_64B descriptor (bf16 tiles of 32 elements = 64-byte rows): 8-row groups 512 B apart, layout type 4
__device__ __forceinline__ uint64_t make_kmajor_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::f16 (bf16 operands, fp32 accumulator), K = 16 per instruction
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {      // round to nearest even, lo in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// The landed fp32 tile [rows x 32] (128-byte rows, 128B swizzle: 16-byte chunk j of row r at j ^ (r % 8)) -> bf16 tile
// [rows x 32] at the same base (64-byte rows, 64B swizzle: 16-byte chunk c of row r at c ^ ((r / 2) % 4)).  In place:
// every thread reads its items, ALL threads pass `sync`, then they write.  item = (row, bf16 chunk c = 8 elements).
template <int NTHREADS, typename Sync>
__device__ __forceinline__ void tile_to_bf16_inplace(uint32_t tile, int rows, uint32_t t, Sync sync) {
    constexpr int kMaxItems = (128 * 4 + NTHREADS - 1) / NTHREADS;
    float4 a[kMaxItems], b[kMaxItems];
#pragma unroll
    for (int i = 0; i < kMaxItems; ++i) {
        const uint32_t item = t + (uint32_t)i * NTHREADS;
        const uint32_t r = item >> 2, c = item & 3u;
        if ((int)r < rows) {
            const uint32_t base = tile + r * 128u, sw = r & 7u;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(a[i].x), "=f"(a[i].y), "=f"(a[i].z), "=f"(a[i].w) : "r"(base + (((2u * c) ^ sw) << 4)));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b[i].x), "=f"(b[i].y), "=f"(b[i].z), "=f"(b[i].w) : "r"(base + (((2u * c + 1u) ^ sw) << 4)));
        }
    }
    sync();
#pragma unroll
    for (int i = 0; i < kMaxItems; ++i) {
        const uint32_t item = t + (uint32_t)i * NTHREADS;
        const uint32_t r = item >> 2, c = item & 3u;
        if ((int)r < rows)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(tile + r * 64u + ((c ^ ((r >> 1) & 3u)) << 4)), "r"(pack_bf16(a[i].x, a[i].y)),
                           "r"(pack_bf16(a[i].z, a[i].w)), "r"(pack_bf16(b[i].x, b[i].y)), "r"(pack_bf16(b[i].z, b[i].w))
                         : "memory");
    }
}

template <int PASSES>
__global__ void __launch_bounds__(kGemmThreads, 1)
tf32_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_b_lo, float* __restrict__ C, int M, int N, int K, int n_tile,
                 int n_tiles, int tmem_cols, int STAGES, const MaskEpilogue ep, int kb_per_split, int sm_count) {
    GPT_PDL_TRIGGER();
    if (ep.m_live != nullptr) {     // CTA-uniform, before any barrier or tensor-memory state exists
        GPT_PDL_WAIT();
        const int m_live = *ep.m_live;
        if ((int)(blockIdx.x / n_tiles) * BM >= m_live) return;
        if (gridDim.y > 1) {
            // split-K with a device-side row count: the host launched the finest split it allows (gridDim.y ranges); the
            // number actually used is chosen here, so that the live row tiles x splits fill the SMs once -- every CTA
            // streams its K range through one SM's L2 port, so idle SMs are lost bandwidth -- but never more k-blocks per
            // accumulator chain than kMaxChainKb
            const int nkb_all = (K + BK - 1) / BK;
            const int live_tiles = ((m_live + BM - 1) / BM) * n_tiles;
            int splits = sm_count / live_tiles;
            splits = max(splits, (nkb_all + kMaxChainKb - 1) / kMaxChainKb);
            splits = max(1, min(splits, (int)gridDim.y));
            kb_per_split = (nkb_all + splits - 1) / splits;
            if ((int)blockIdx.y * kb_per_split >= nkb_all) return;
        }
    }
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[3 * kMaxStages + 1];
    __shared__ uint32_t tmem_base_holder;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // N tiles of one M tile are neighbours in launch order: they run at the same time and the second read of the A tile
    // is served by L2 instead of HBM
    const int m0 = (int)(blockIdx.x / n_tiles) * BM, n0 = (int)(blockIdx.x % n_tiles) * n_tile;
    // split-K (blockIdx.y): a long reduction over few output tiles (the relation-aware layers' data gradient: 3 000 x 200
    // outputs reduced over D*H = 10 000) is cut into ranges of kb_per_split k-blocks; the partial tiles meet in C through
    // vector reductions (the launcher zeroes C first)
    const int nkb_all = (K + BK - 1) / BK;
    const int kb0 = (int)blockIdx.y * kb_per_split;
    const int nkb = min(kb_per_split, nkb_all - kb0);
    const bool split_k = gridDim.y > 1;
    const uint32_t a_bytes = BM * BK * 4, b_bytes = (uint32_t)n_tile * BK * (PASSES == 0 ? 2 : 4);
    // stage layout: [A | A_lo | B | B_lo] (the lo tiles only for PASSES == 3); every tile is 1024-byte aligned
    const uint32_t stage_bytes = (PASSES == 3 ? 2u : 1u) * (a_bytes + b_bytes);
    const uint32_t off_alo = a_bytes, off_b = (PASSES == 3 ? 2u : 1u) * a_bytes, off_blo = off_b + b_bytes;
    const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-byte aligned tiles
    const uint32_t full0 = smem_addr(&bars[0]), empty0 = smem_addr(&bars[kMaxStages]);
    const uint32_t split0 = smem_addr(&bars[2 * kMaxStages]), done = smem_addr(&bars[3 * kMaxStages]);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
        if (PASSES == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b_lo)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
            mbar_init(split0 + 8 * s, 32 * kSplitWarps);   // every splitter thread arrives
        }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // one warp allocates the accumulator columns and owns the deallocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_addr(&tmem_base_holder)), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_holder;
    GPT_PDL_WAIT();     // barriers, tensor memory and descriptors are ready; only now are the operands (and C) touched

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(empty0 + 8 * s, ((kb / STAGES) - 1) & 1);
                const uint32_t st = tiles + (uint32_t)s * stage_bytes;
                mbar_expect_tx(full0 + 8 * s, a_bytes + (PASSES == 3 ? 2u : 1u) * b_bytes);
                tma_load_2d(st, &tm_a, full0 + 8 * s, (kb0 + kb) * BK, m0);      // OOB rows / columns arrive as zeros
                tma_load_2d(st + off_b, &tm_b, full0 + 8 * s, (kb0 + kb) * BK, n0);
                if (PASSES == 3) tma_load_2d(st + off_blo, &tm_b_lo, full0 + 8 * s, (kb0 + kb) * BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9,
            // 10-12 = 2), both K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28
            // kind::f16 with bf16 operands: A = B = BF16 (format 1)
            const uint32_t fmt = PASSES == 0 ? 1u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n_tile >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                // PASSES == 3 / 0: the splitters arrive after the TMA bytes have landed and A_lo / bf16(A) is written
                mbar_wait((PASSES != 1 ? split0 : full0) + 8 * s, (kb / STAGES) & 1);
                tc_fence_after();
                const uint32_t st = tiles + (uint32_t)s * stage_bytes;
                if (PASSES == 0) {
                    const uint64_t a16 = make_kmajor_desc_sw64(st), b16 = make_kmajor_desc_sw64(st + off_b);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)      // 16 bf16 = 32 bytes per instruction
                        umma_bf16(tmem_base, a16 + 2 * k, b16 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty0 + 8 * s);
                    continue;
                }
                const uint64_t a_desc = make_kmajor_desc(st), b_desc = make_kmajor_desc(st + off_b);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)  // +32 bytes inside the swizzle atom = +2 in the address field
                    umma_tf32(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                if (PASSES == 3) {
                    const uint64_t alo_desc = make_kmajor_desc(st + off_alo), blo_desc = make_kmajor_desc(st + off_blo);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) umma_tf32(tmem_base, alo_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) umma_tf32(tmem_base, a_desc + 2 * k, blo_desc + 2 * k, idesc, 1u);
                }
                umma_commit(empty0 + 8 * s);           // slot is free once these MMAs have read it
            }
            umma_commit(done);                          // accumulator complete
        }
    } else {
        if (PASSES == 0) {
            // ===== converters: the landed fp32 X tile -> bf16 (in place) =====
            const uint32_t t = threadIdx.x - 64;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
                tile_to_bf16_inplace<32 * kSplitWarps>(tiles + (uint32_t)s * stage_bytes, BM, t, [] {
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSplitWarps) : "memory");
                });
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(split0 + 8 * s);
            }
        }
        if (PASSES == 3) {
            // ===== splitters: A -> A_hi (in place), A_lo = A - A_hi, element-wise in the swizzled tile =====
            const uint32_t t = threadIdx.x - 64;       // 0..255: 16-byte chunk t, t + 256, ... of the A tile
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
                const uint32_t src = tiles + (uint32_t)s * stage_bytes + t * 16u, dst = src + off_alo;
                constexpr uint32_t kIters = (BM * BK * 4) / (32 * kSplitWarps * 16);
                float4 v[kIters];
#pragma unroll
                for (uint32_t i = 0; i < kIters; ++i)       // all loads of the thread in flight together
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                                 : "r"(src + i * (32u * kSplitWarps * 16u)));
#pragma unroll
                for (uint32_t i = 0; i < kIters; ++i) {
                    const float4 h = make_float4(tf32_hi(v[i].x), tf32_hi(v[i].y), tf32_hi(v[i].z), tf32_hi(v[i].w));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(src + i * (32u * kSplitWarps * 16u)),
                                 "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + i * (32u * kSplitWarps * 16u)),
                                 "f"(v[i].x - h.x), "f"(v[i].y - h.y), "f"(v[i].z - h.z), "f"(v[i].w - h.w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core
                mbar_arrive(split0 + 8 * s);
            }
        }
        // ===== epilogue: TMEM -> registers -> global =====
        mbar_wait(done, 0);
        tc_fence_after();
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        float* crow = C + (size_t)row * N + n0;
        const bool vec_ok = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        float ep_inv = 0.f;
        const uint32_t* ep_words = nullptr;                 // act word of (sentence, 32-column block 0, token)
        if (ep.act != nullptr && row < M) {
            const int bb = row / ep.T, tt = row - bb * ep.T;
            ep_inv = __frcp_rn(ep.denom[row]);
            ep_words = ep.act + ((size_t)bb * ((N + 31) / 32)) * ep.T + tt;
        }
        // warps 2..5 take the first half of the tile's 32-column blocks, warps 6..9 the second half
        const int nblk = (n_tile + 31) / 32, half_blk = (nblk + 1) / 2;
        const int c_begin = (warp - 2) < 4 ? 0 : half_blk * 32, c_end = (warp - 2) < 4 ? min(n_tile, half_blk * 32) : n_tile;
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                  "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                  "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ep.bias != nullptr) {            // N % 4 == 0 and n0 + c0 is a multiple of 16: aligned 128-bit loads
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int c = n0 + c0 + j;
                    if (c + 3 < N) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + c));
                        v[j] = __float_as_uint(__uint_as_float(v[j]) + b.x);
                        v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
                        v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
                        v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
                    }
                }
            }
            if (ep_words != nullptr) {                      // g = (dx * (bit * scale)) * (1 / denom), as K2 forms it
                const int cg = n0 + c0, wi = cg >> 5, sh = cg & 31, nw = (N + 31) / 32;
                unsigned long long bits = wi < nw ? (unsigned long long)ep_words[(size_t)wi * ep.T] : 0ull;
                if (sh != 0 && wi + 1 < nw) bits |= (unsigned long long)ep_words[(size_t)(wi + 1) * ep.T] << 32;
                const uint32_t m = (uint32_t)(bits >> sh);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float f = (float)((m >> j) & 1u) * ep.scale;
                    v[j] = __float_as_uint(__uint_as_float(v[j]) * f * ep_inv);
                }
            }
            if (row < M) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int c = n0 + c0 + j;
                    if (c0 + j >= n_tile) break;        // n_tile is a multiple of 16, not of 32: stay inside this tile
                    if (split_k) {
                        if (vec_ok && c + 3 < N) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + c0 + j),
                                         "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                         "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (c + i < N) atomicAdd(crow + c0 + j + i, __uint_as_float(v[j + i]));
                        }
                    } else if (vec_ok && c + 3 < N) {
                        *reinterpret_cast<float4*>(crow + c0 + j) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                        __uint_as_float(v[j + 3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (c + i < N) crow[c0 + j + i] = __uint_as_float(v[j + i]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols)
                     : "memory");
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float t[32][33];
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        if (x < cols && y0 + j < rows) t[j][threadIdx.x] = in[(size_t)(y0 + j) * cols + x];
    __syncthreads();
    const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        if (ox < rows && oy0 + j < cols) out[(size_t)(oy0 + j) * rows + ox] = t[threadIdx.x][j];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// row-major fp32 [rows, cols] -> boxes of [box_rows x 32] floats, 128-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap* map, const float* base, int rows, int cols, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return GPT_ERR_DRIVER;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GPT_OK : GPT_ERR_DRIVER;
}

// row-major bf16 [rows, cols] -> boxes of [box_rows x 32] elements (64-byte rows), 64-byte swizzle, zero fill
int make_map_bf16(CUtensorMap* map, const void* base, int rows, int cols, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return GPT_ERR_DRIVER;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GPT_OK : GPT_ERR_DRIVER;
}

__global__ void tf32_split_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo,
                                  size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = in[i], h = tf32_hi(v);
        hi[i] = h;
        lo[i] = v - h;
    }
}

#include "gemm_persist.cuh"   // the large-M path: persistent CTA-pair kernel (uses the helpers above)

inline int sm_count_host() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <int PASSES>
int launch_gemm(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_b_lo, float* C, int M, int N,
                int K, int n_tile, int n_tiles, int tmem_cols, cudaStream_t st, const MaskEpilogue& ep, int splits = 1) {
    const size_t stage = (size_t)(PASSES == 3 ? 2 : 1) * (BM * BK * 4 + (size_t)n_tile * BK * (PASSES == 0 ? 2 : 4));
    int stages = (int)((200 * 1024) / stage);
    const int nkb_all = (K + BK - 1) / BK;
    int kb_per_split = (nkb_all + splits - 1) / splits;
    splits = (nkb_all + kb_per_split - 1) / kb_per_split;
    const int nkb = kb_per_split;
    stages = stages > kMaxStages ? kMaxStages : stages;
    stages = stages > nkb ? nkb : stages;
    stages = stages < 1 ? 1 : stages;
    const size_t smem = (size_t)stages * stage + 1024;
    if (int a = gpt_smem_opt_in(tf32_gemm_kernel<PASSES>, smem)) return a;
    if (splits > 1 && ep.bias != nullptr) return GPT_ERR_UNSUPPORTED;     // every K range would add it
    if (splits > 1) {               // the partial tiles are ADDED into C
        const cudaError_t e = cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    const int sms = sm_count_host();
    dim3 grid((unsigned)(((M + BM - 1) / BM) * n_tiles), (unsigned)splits);
    gpt_launch(tf32_gemm_kernel<PASSES>, grid, dim3(kGemmThreads), smem, st, tm_a, tm_b, tm_b_lo, C, M, N, K, n_tile, n_tiles,
                                                               tmem_cols, stages, ep, kb_per_split, sms);
    return gpt_launch_status();
}

// C[M,N] = A[M,K] . B[N,K]^T ; b_lo != nullptr selects the 3xTF32 mode (B = rounded hi part, b_lo = lo part)
int run_tf32_gemm(const float* A, const float* B, const float* b_lo, float* C, int M, int N, int K, cudaStream_t st,
                  const MaskEpilogue ep = MaskEpilogue{nullptr, nullptr, 1.f, 1}) {
    if (M == 0) return GPT_OK;
    // TMA: 16-byte aligned bases and row pitches
    if (K % 4 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
        (reinterpret_cast<uintptr_t>(b_lo) & 15))
        return GPT_ERR_UNSUPPORTED;
    {   // large M: persistent CTA-pair kernel with double-buffered accumulators and a TMA-store epilogue
        const int rc = persist::run(A, B, b_lo, C, M, N, K, st, ep);
        if (rc != GPT_ERR_UNSUPPORTED) return rc;
    }
    // N tile: as wide as possible (A is re-read once per N tile) unless that leaves most SMs idle; a narrower tile
    // also makes the stages smaller, i.e. the ring deeper, which is what hides TMA latency when M is small
    const int m_tiles = (M + BM - 1) / BM;
    const int nkb_all = (K + BK - 1) / BK;
    int cap = 256, splits = 1;
    if (nkb_all >= 64 && (ep.m_live != nullptr || (long)m_tiles * ((N + 255) / 256) < 74)) {
        // long reduction, few output tiles: keep the N tile wide (the X tile is loaded and split once per N tile) and
        // fill the machine by splitting K instead
        splits = 148 / (m_tiles * ((N + 255) / 256));
        if (splits > nkb_all / 8) splits = nkb_all / 8;
        // the tensor core accumulates round-toward-zero (~1.8e-8 relative per addition, wgrad_tcgen05.cu): at most
        // kMaxChainKb k-blocks (x 4 K-steps x 3 passes = 384 additions, <= 7e-6) go into one accumulator chain; a second
        // wave of CTAs costs less than the fp32 grade of the result
        if (splits < (nkb_all + kMaxChainKb - 1) / kMaxChainKb) splits = (nkb_all + kMaxChainKb - 1) / kMaxChainKb;
        // device-side row count: how many row tiles are live is not known here; launch the finest split (8 k-blocks per
        // range) and let the kernel choose (tf32_gemm_kernel)
        if (ep.m_live != nullptr) splits = nkb_all / 8;
        if (splits < 1) splits = 1;
    } else {
        // halve the tile while the grid still fits ONE wave (one CTA per SM at these shared-memory sizes): 168 tiles on 148
        // SMs run as two waves -- the first layer's data gradient (N = 360) took 18.8 us that way against 13.3 for 112 tiles
        while (cap > 64 && (long)m_tiles * ((N + cap / 2 - 1) / (cap / 2)) <= sm_count_host()) cap >>= 1;
    }
    const int n_tiles = (N + cap - 1) / cap;
    int n_tile = ((N + n_tiles - 1) / n_tiles + 15) / 16 * 16;   // UMMA N: multiple of 16 at M = 128, <= 256
    if (n_tile < 16) n_tile = 16;
    int tmem_cols = 32;
    while (tmem_cols < n_tile) tmem_cols <<= 1;
    alignas(64) CUtensorMap tm_a, tm_b, tm_b_lo;
    int rc = make_map(&tm_a, A, M, K, BM);
    if (rc != GPT_OK) return rc;
    if ((rc = make_map(&tm_b, B, N, K, n_tile)) != GPT_OK) return rc;
    if ((rc = make_map(&tm_b_lo, b_lo ? b_lo : B, N, K, n_tile)) != GPT_OK) return rc;
    if (b_lo != nullptr) return launch_gemm<3>(tm_a, tm_b, tm_b_lo, C, M, N, K, n_tile, n_tiles, tmem_cols, st, ep, splits);
    return launch_gemm<1>(tm_a, tm_b, tm_b_lo, C, M, N, K, n_tile, n_tiles, tmem_cols, st, ep, splits);
}

// w [N,K] -> ws = [w_hi | w_lo | wt_hi | wt_lo]  (wt = w^T [K,N]); hi = round_tf32, lo = w - hi
__global__ void weight_prep_kernel(const float* __restrict__ w, float* __restrict__ ws, int N, int K) {
    __shared__ float th[32][33], tl[32][33];
    const size_t nk = (size_t)N * K;
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;   // x: k index, y: n index
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        if (x < K && y0 + j < N) {
            const size_t i = (size_t)(y0 + j) * K + x;
            const float v = w[i], h = tf32_hi(v);
            ws[i] = h;
            ws[nk + i] = v - h;
            th[j][threadIdx.x] = h;
            tl[j][threadIdx.x] = v - h;
        }
    }
    __syncthreads();
    const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;  // ox: n index, oy: k index
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        if (ox < N && oy0 + j < K) {
            const size_t o = (size_t)(oy0 + j) * N + ox;
            ws[2 * nk + o] = th[threadIdx.x][j];
            ws[3 * nk + o] = tl[threadIdx.x][j];
        }
    }
}

// every layer's weight in one launch (blockIdx.z = layer): the side stream that prepares the operands then needs one
// launch latency instead of one per layer before the first projection may start
struct PrepBatch {
    const float* w[8];
    float* ws[8];
    int N[8], K[8];
};
__global__ void weight_prep_batch_kernel(const PrepBatch b) {
    __shared__ float th[32][33], tl[32][33];
    const int l = blockIdx.z;
    const float* __restrict__ w = b.w[l];
    float* __restrict__ ws = b.ws[l];
    const int N = b.N[l], K = b.K[l];
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    if (blockIdx.x * 32 >= K || y0 >= N) return;              // (CTA-uniform) this layer's matrix is smaller than the grid
    const size_t nk = (size_t)N * K;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        if (x < K && y0 + j < N) {
            const size_t i = (size_t)(y0 + j) * K + x;
            const float v = w[i], h = tf32_hi(v);
            ws[i] = h;
            ws[nk + i] = v - h;
            th[j][threadIdx.x] = h;
            tl[j][threadIdx.x] = v - h;
        }
    }
    __syncthreads();
    const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        if (ox < N && oy0 + j < K) {
            const size_t o = (size_t)(oy0 + j) * N + ox;
            ws[2 * nk + o] = th[threadIdx.x][j];
            ws[3 * nk + o] = tl[threadIdx.x][j];
        }
    }
}

int split_hi_lo(const float* in, float* hi, float* lo, size_t n, cudaStream_t st) {  // hi may alias in
    tf32_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, hi, lo, n);
    return gpt_launch_status();
}

int transpose(const float* w, float* wt, int N, int K, cudaStream_t st) {
    transpose_kernel<<<dim3((K + 31) / 32, (N + 31) / 32), dim3(32, 8), 0, st>>>(w, wt, N, K);
    return gpt_launch_status();
}

}  // namespace

namespace {

// C[M,N] = bf16(A[M,K]) . B16[N,K]^T with B16 already bf16 (gpt_weight_prep_bf16); fp32 accumulation
int run_bf16_gemm(const float* A, const void* B16, float* C, int M, int N, int K, cudaStream_t st) {
    if (M == 0) return GPT_OK;
    if (K % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B16) & 15))
        return GPT_ERR_UNSUPPORTED;              // TMA: 16-byte aligned bases and row pitches (bf16 rows: K % 8)
    const MaskEpilogue ep{nullptr, nullptr, 1.f, 1};
    {
        const int rc = persist::run_bf16(A, B16, C, M, N, K, st, ep);
        if (rc != GPT_ERR_UNSUPPORTED) return rc;
    }
    const int m_tiles = (M + BM - 1) / BM;
    int cap = 256;
    while (cap > 64 && (long)m_tiles * ((N + cap / 2 - 1) / (cap / 2)) <= sm_count_host()) cap >>= 1;
    const int n_tiles = (N + cap - 1) / cap;
    int n_tile = ((N + n_tiles - 1) / n_tiles + 15) / 16 * 16;
    if (n_tile < 16) n_tile = 16;
    int tmem_cols = 32;
    while (tmem_cols < n_tile) tmem_cols <<= 1;
    alignas(64) CUtensorMap tm_a, tm_b;
    int rc = make_map(&tm_a, A, M, K, BM);
    if (rc != GPT_OK) return rc;
    if ((rc = make_map_bf16(&tm_b, B16, N, K, n_tile)) != GPT_OK) return rc;
    return launch_gemm<0>(tm_a, tm_b, tm_b, C, M, N, K, n_tile, n_tiles, tmem_cols, st, ep);
}

// w [N,K] fp32 -> ws16 = [ bf16(w) [N,K] | bf16(w^T) [K,N] ]  (2*N*K bf16 = N*K floats of workspace)
__global__ void weight_prep_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ ws, int N, int K) {
    __shared__ float t[32][33];
    const size_t nk = (size_t)N * K;
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        if (x < K && y0 + j < N) {
            const float v = w[(size_t)(y0 + j) * K + x];
            ws[(size_t)(y0 + j) * K + x] = __float2bfloat16_rn(v);
            t[j][threadIdx.x] = v;
        }
    }
    __syncthreads();
    const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        if (ox < N && oy0 + j < K) ws[nk + (size_t)(oy0 + j) * N + ox] = __float2bfloat16_rn(t[threadIdx.x][j]);
}

}  // namespace

extern "C" int gpt_weight_prep_bf16(const float* w, void* ws, int N, int K, void* stream) {
    GPT_CHECK_ARG(w && ws && N >= 1 && K >= 1);
    weight_prep_bf16_kernel<<<dim3((K + 31) / 32, (N + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        w, reinterpret_cast<__nv_bfloat16*>(ws), N, K);
    return gpt_launch_status();
}

extern "C" int gpt_linear_fwd_bf16(const float* x, const void* ws, float* y, int M, int N, int K, void* stream) {
    GPT_CHECK_ARG(x && ws && y && M >= 0 && N >= 1 && K >= 1);
    return run_bf16_gemm(x, ws, y, M, N, K, (cudaStream_t)stream);
}

extern "C" int gpt_linear_dgrad_bf16(const float* dy, const void* ws, float* dx, int M, int N, int K, void* stream) {
    GPT_CHECK_ARG(dy && ws && dx && M >= 0 && N >= 1 && K >= 1);
    const __nv_bfloat16* wt = reinterpret_cast<const __nv_bfloat16*>(ws) + (size_t)N * K;     // bf16(w^T) [K, N]
    return run_bf16_gemm(dy, wt, dx, M, K, N, (cudaStream_t)stream);
}

extern "C" int gpt_gemm_persist_config(int cta_group, long long min_rows) {
    GPT_CHECK_ARG(cta_group >= 0 && cta_group <= 2 && min_rows >= 0);
    persist::config() = persist::Config{cta_group, min_rows};
    return GPT_OK;
}

extern "C" int gpt_linear_fwd_tf32(const float* x, const float* w, float* y, int M, int N, int K, void* stream) {
    GPT_CHECK_ARG(x && w && y && M >= 0 && N >= 1 && K >= 1);
    return run_tf32_gemm(x, w, nullptr, y, M, N, K, (cudaStream_t)stream);
}

extern "C" int gpt_linear_dgrad_tf32(const float* dy, const float* w, float* dx, float* workspace, int M, int N, int K,
                                     void* stream) {
    GPT_CHECK_ARG(dy && w && dx && workspace && M >= 0 && N >= 1 && K >= 1);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = transpose(w, workspace, N, K, st);  // operand B must be K-major over the reduction index n: W^T [K, N]
    if (rc != GPT_OK) return rc;
    return run_tf32_gemm(dy, workspace, nullptr, dx, M, K, N, st);
}

extern "C" int gpt_weight_prep_tf32x3(const float* w, float* ws, int N, int K, void* stream) {
    GPT_CHECK_ARG(w && ws && N >= 1 && K >= 1);
    weight_prep_kernel<<<dim3((K + 31) / 32, (N + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(w, ws, N, K);
    return gpt_launch_status();
}

extern "C" int gpt_weight_prep_tf32x3_batch(const float* const* w, float* const* ws, const int* N, const int* K,
                                            int n_layers, void* stream) {
    GPT_CHECK_ARG(w && ws && N && K && n_layers >= 1 && n_layers <= 8);
    PrepBatch b{};
    int max_n = 0, max_k = 0;
    for (int l = 0; l < n_layers; ++l) {
        GPT_CHECK_ARG(w[l] && ws[l] && N[l] >= 1 && K[l] >= 1);
        b.w[l] = w[l]; b.ws[l] = ws[l]; b.N[l] = N[l]; b.K[l] = K[l];
        max_n = N[l] > max_n ? N[l] : max_n;
        max_k = K[l] > max_k ? K[l] : max_k;
    }
    weight_prep_batch_kernel<<<dim3((max_k + 31) / 32, (max_n + 31) / 32, n_layers), dim3(32, 8), 0,
                               (cudaStream_t)stream>>>(b);
    return gpt_launch_status();
}

extern "C" int gpt_linear_fwd_tf32x3(const float* x, const float* ws, float* y, int M, int N, int K, void* stream) {
    GPT_CHECK_ARG(x && ws && y && M >= 0 && N >= 1 && K >= 1);
    return run_tf32_gemm(x, ws, ws + (size_t)N * K, y, M, N, K, (cudaStream_t)stream);
}

extern "C" int gpt_linear_dgrad_tf32x3(const float* dy, const float* ws, float* dx, int M, int N, int K,
                                       void* stream) {
    GPT_CHECK_ARG(dy && ws && dx && M >= 0 && N >= 1 && K >= 1);
    const size_t nk = (size_t)N * K;
    return run_tf32_gemm(dy, ws + 2 * nk, ws + 3 * nk, dx, M, K, N, (cudaStream_t)stream);
}

// the same two projections over the first *m_live rows only (device-side count, gpt_live_rows); rows beyond it are not
// computed and the corresponding rows of the output keep whatever they held
extern "C" int gpt_linear_fwd_tf32x3_rows(const float* x, const float* ws, const float* bias, float* y, int M, int N,
                                          int K, const int32_t* m_live, void* stream) {
    GPT_CHECK_ARG(x && ws && y && m_live && M >= 0 && N >= 1 && K >= 1);
    if (bias != nullptr && (N % 4 != 0 || (reinterpret_cast<uintptr_t>(bias) & 15))) return GPT_ERR_UNSUPPORTED;
    return run_tf32_gemm(x, ws, ws + (size_t)N * K, y, M, N, K, (cudaStream_t)stream,
                         MaskEpilogue{nullptr, nullptr, 1.f, 1, m_live, bias});
}

extern "C" int gpt_linear_dgrad_tf32x3_rows(const float* dy, const float* ws, float* dx, int M, int N, int K,
                                            const int32_t* m_live, void* stream) {
    GPT_CHECK_ARG(dy && ws && dx && m_live && M >= 0 && N >= 1 && K >= 1);
    const size_t nk = (size_t)N * K;
    return run_tf32_gemm(dy, ws + 2 * nk, ws + 3 * nk, dx, M, K, N, (cudaStream_t)stream,
                         MaskEpilogue{nullptr, nullptr, 1.f, 1, m_live, nullptr});
}

extern "C" int gpt_linear_dgrad_tf32x3_masked(const float* dy, const float* ws, float* g, const uint32_t* act_prev,
                                              const float* denom, float drop_scale_prev, int T, int M, int N, int K,
                                              void* stream) {
    GPT_CHECK_ARG(dy && ws && g && act_prev && denom && T >= 1 && M >= 0 && M % T == 0 && N >= 1 && K >= 1);
    const size_t nk = (size_t)N * K;
    return run_tf32_gemm(dy, ws + 2 * nk, ws + 3 * nk, g, M, K, N, (cudaStream_t)stream,
                         MaskEpilogue{act_prev, denom, drop_scale_prev, T});
}
