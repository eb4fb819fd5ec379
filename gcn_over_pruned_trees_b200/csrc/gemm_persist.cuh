// K3 at scale -- persistent CTA-pair tcgen05 GEMM (sm_100a), included by gemm_tcgen05.cu.
//
//   Y[M,N] = X[M,K] . W[N,K]^T     fp32 in HBM, TF32 (1 pass) or 3xTF32 (fp32-grade) on the tensor cores
//
// The one-tile-per-CTA kernel in gemm_tcgen05.cu is right for TACRED-sized batches (a few dozen tiles, latency-bound).
// At the large shape (BASELINE.json configs[4]: 2 097 152 rows x 512 x {360, 512}) ncu showed it at 26 % / 47 % tensor-pipe
// activity: prologue, pipeline fill and a strided register -> global epilogue were never overlapped with MMAs, and every
// 128-row tile re-read the whole of W from L2.  This kernel is the large-shape path:
//
//   * persistent: one cluster per SM pair, static schedule over 256-row blocks; the N tiles of a block are consecutive on
//     the same cluster, so the second read of the block's X rows is an L2 hit by construction
//   * CTA pair (CG = 2): tcgen05.mma.cta_group::2 with M = 256 -- each CTA stages its own 128 rows of X and HALF of the W
//     tile, the tensor cores of both SMs read both halves: W's L2 -> SMEM traffic per output row is halved
//   * two accumulators in tensor memory (2 x 256 columns): the epilogue of tile i runs while the MMAs of tile i+1 issue
//   * epilogue through shared memory and TMA: tcgen05.ld -> registers -> 128B-swizzled staging tile ->
//     cp.async.bulk.tensor store (coalesced 16 KB boxes, no per-thread global stores, bounds clipped by the descriptor)
//   * 3xTF32: dedicated splitter warps rewrite the landed X tile as hi and write lo next to it (as in the small kernel)
//
// Warp roles (448 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer (leader CTA only issues), 2..5 = epilogue
// (TMEM lane quarter = warp % 4), 6..13 = operand splitters (3xTF32) / relay (TF32).
//
// Barriers (all in each CTA's shared memory; "leader" = cluster rank 0):
//   full[s]       TMA bytes of this CTA's stage s have landed                      (local, tx-count)
//   ready[s]      stage s is consumable in BOTH CTAs: one arrival per CTA on the LEADER's barrier, sent by the splitters
//                 (after hi/lo are written and fenced to the async proxy) or by the relay thread
//   empty[s]      the MMAs reading stage s have completed: tcgen05.commit multicast to both CTAs
//   tmem_full[a]  accumulator a is complete: tcgen05.commit multicast to both CTAs
//   tmem_empty[a] accumulator a has been read out: one arrival per epilogue warp of both CTAs on the LEADER's barrier
#pragma once

namespace persist {

constexpr int BM = 128;                 // rows per CTA (UMMA M = 128 * CG)
constexpr int BK = 32;                  // fp32 per k-block = one 128-byte swizzle atom
constexpr int UK = 8;                   // kind::tf32: 8 elements of K per instruction
constexpr int kEpiWarps = 4;
constexpr int kSplitWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps + kSplitWarps);     // 448
constexpr int kMaxStages = 8;
constexpr int kAccCols = 256;           // columns per accumulator stage (2 stages = the SM's whole tensor memory)
constexpr uint32_t kEpiBuf = BM * 32 * 4;   // staging tile of the TMA store: 128 rows x 32 fp32 = 16 KB

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t n_clusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory variable in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier whose phase is completed by arrivals from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
template <int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (CG == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
            : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void umma16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (CG == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
            : "memory");
    }
}
// arrive (once every MMA issued so far by this thread has completed) on the barrier at this offset in every CTA of the pair
template <int CG>
__device__ __forceinline__ void commit_all(uint32_t bar) {
    if (CG == 2) {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"((unsigned short)3)
            : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_all(uint32_t holder, uint32_t cols) {
    if (CG == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_all(uint32_t base, uint32_t cols) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the PEER CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void named_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Static tile schedule of one cluster: its q-th tile.  Many row blocks (the large shapes): a cluster owns whole row blocks
// and walks their N tiles back to back, so the second read of the block's X rows is an L2 hit.  Few row blocks and many N
// tiles (the relation-aware layers: ~1 000 live rows x D*H = 10 000 columns): tiles are dealt round-robin, or most
// clusters would have nothing to do.
struct TileSched {
    int m_blocks, n_tiles, cluster, n_clusters, flat;
    __device__ __forceinline__ bool get(int q, int& mb, int& nb) const {
        if (flat) {
            const int t = cluster + q * n_clusters;
            mb = t / n_tiles;
            nb = t - mb * n_tiles;
            return t < m_blocks * n_tiles;
        }
        mb = cluster + (q / n_tiles) * n_clusters;
        nb = q % n_tiles;
        return mb < m_blocks;
    }
};

template <int PASSES, int CG>
__global__ void __launch_bounds__(kThreads, 1)
gemm_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       const __grid_constant__ CUtensorMap tm_b_lo, const __grid_constant__ CUtensorMap tm_c,
                       int M, int N, int K, int n_tile, int n_tiles, int STAGES, const MaskEpilogue ep, int flat) {
    GPT_PDL_TRIGGER();
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[3 * kMaxStages + 4];
    __shared__ uint32_t tmem_base_holder;
    __shared__ __align__(16) float s_bias[kAccCols];    // the tile's column bias (one copy: the epilogue's own barriers order it)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
    const bool leader = rank == 0;
    const int cluster = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
    const int n_clusters = CG == 2 ? (int)n_clusters_x() : (int)gridDim.x;
    const int nkb = (K + BK - 1) / BK;
    const int n_half = n_tile / CG;                                 // rows of W this CTA stages per tile
    const uint32_t a_bytes = BM * BK * 4, b_bytes = (uint32_t)n_half * BK * (PASSES == 0 ? 2 : 4);
    constexpr uint32_t kLo = PASSES == 3 ? 2u : 1u;
    constexpr bool kDirect = PASSES == 1;       // TF32: the MMA thread waits on the TMA barrier itself (no splitters, no relay)
    const uint32_t stage_bytes = kLo * (a_bytes + b_bytes);         // [A | A_lo | B | B_lo], every tile 1024-byte aligned
    const uint32_t off_alo = a_bytes, off_b = kLo * a_bytes, off_blo = off_b + b_bytes;
    const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_buf = tiles + (uint32_t)STAGES * stage_bytes;            // 2 x 16 KB, 1024-byte aligned
    const uint32_t full0 = smem_addr(&bars[0]), empty0 = smem_addr(&bars[kMaxStages]);
    const uint32_t ready0 = smem_addr(&bars[2 * kMaxStages]);
    const uint32_t tfull0 = smem_addr(&bars[3 * kMaxStages]), tempty0 = smem_addr(&bars[3 * kMaxStages + 2]);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_c)) : "memory");
        if (PASSES == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b_lo)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
            mbar_init(ready0 + 8 * s, CG);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, CG * kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_all<CG>(smem_addr(&tmem_base_holder), 2 * kAccCols);
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();        // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_holder;
    GPT_PDL_WAIT();
    // the row count may live on the device (gpt_live_rows): row blocks beyond it are not scheduled at all
    const int m_rows = ep.m_live != nullptr ? min(M, *ep.m_live) : M;
    const TileSched sched{(m_rows + CG * BM - 1) / (CG * BM), n_tiles, cluster, n_clusters, flat};
    int mb, nb;

    if (warp == 0) {
        // ===== TMA producer (each CTA: its 128 rows of X, its half of the W tile) =====
        if (lane == 0) {
            uint32_t it = 0;
            // TF32 on a CTA pair: nothing has to touch the landed tiles, so BOTH CTAs' loads credit their bytes to the
            // leader's full[s] (cp.async.bulk.tensor.cta_group::2) and the MMA thread waits on that barrier directly
            const uint32_t full_leader = (kDirect && CG == 2) ? map_to_rank(full0, 0) : full0;
            for (int q = 0; sched.get(q, mb, nb); ++q) {
                const int m0 = (mb * CG + (int)rank) * BM;
                const int n0 = nb * n_tile + (int)rank * n_half;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % (uint32_t)STAGES;
                    if (it >= (uint32_t)STAGES) mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
                    const uint32_t st = tiles + s * stage_bytes;
                    if (kDirect && CG == 2) {
                        if (leader) mbar_expect_tx(full0 + 8 * s, 2u * (a_bytes + b_bytes));
                        tma_load_2d_pair(st, &tm_a, full_leader + 8 * s, kb * BK, m0);
                        tma_load_2d_pair(st + off_b, &tm_b, full_leader + 8 * s, kb * BK, n0);
                    } else {
                        mbar_expect_tx(full0 + 8 * s, a_bytes + kLo * b_bytes);
                        tma_load_2d(st, &tm_a, full0 + 8 * s, kb * BK, m0);      // OOB rows / columns arrive as zeros
                        tma_load_2d(st + off_b, &tm_b, full0 + 8 * s, kb * BK, n0);
                        if (PASSES == 3) tma_load_2d(st + off_blo, &tm_b_lo, full0 + 8 * s, kb * BK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives the tensor cores of both SMs =====
        if (leader && lane == 0) {
            const uint32_t fmt = PASSES == 0 ? 1u : 2u;          // operand format: BF16 (kind::f16) or TF32
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n_tile >> 3) << 17) |
                                   ((uint32_t)((CG * BM) >> 4) << 24);
            uint32_t it = 0, tile_it = 0;
            {
                for (int q = 0; sched.get(q, mb, nb); ++q, ++tile_it) {
                    const uint32_t a = tile_it & 1u;
                    if (tile_it >= 2) mbar_wait_cluster(tempty0 + 8 * a, ((tile_it >> 1) - 1) & 1);
                    tc_fence_after();
                    const uint32_t acc = tmem_base + a * kAccCols;
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const uint32_t s = it % (uint32_t)STAGES;
                        mbar_wait_cluster((kDirect ? full0 : ready0) + 8 * s, (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t st = tiles + s * stage_bytes;
                        if (PASSES == 0) {
                            const uint64_t a16 = make_kmajor_desc_sw64(st), b16 = make_kmajor_desc_sw64(st + off_b);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma16<CG>(acc, a16 + 2 * k, b16 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                            commit_all<CG>(empty0 + 8 * s);
                            continue;
                        }
                        const uint64_t a_desc = make_kmajor_desc(st), b_desc = make_kmajor_desc(st + off_b);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma<CG>(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (PASSES == 3) {
                            const uint64_t alo = make_kmajor_desc(st + off_alo), blo = make_kmajor_desc(st + off_blo);
#pragma unroll
                            for (int k = 0; k < BK / UK; ++k) umma<CG>(acc, alo + 2 * k, b_desc + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < BK / UK; ++k) umma<CG>(acc, a_desc + 2 * k, blo + 2 * k, idesc, 1u);
                        }
                        commit_all<CG>(empty0 + 8 * s);      // the slot is free (in both CTAs) once these MMAs have read it
                    }
                    commit_all<CG>(tfull0 + 8 * a);           // accumulator a complete: both epilogues may read it
                }
            }
        }
    } else if (warp >= 2 + kEpiWarps) {
        // ===== splitters (3xTF32 only): rewrite the landed X tile as hi, write lo, tell the leader =====
        const uint32_t t = threadIdx.x - 32 * (2 + kEpiWarps);     // 0..255
        const uint32_t ready_leader = CG == 2 ? map_to_rank(ready0, 0) : ready0;
        uint32_t it = 0;
        {
            for (int q = 0; sched.get(q, mb, nb); ++q) {
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % (uint32_t)STAGES;
                    if (PASSES == 0) {          // bf16 operands: round the landed fp32 X tile to bf16, in place
                        mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
                        tile_to_bf16_inplace<32 * kSplitWarps>(tiles + s * stage_bytes, BM, t,
                                                               [] { named_bar(1, 32 * kSplitWarps); });
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        named_bar(1, 32 * kSplitWarps);
                        if (t == 32u * (it % kSplitWarps)) {
                            if (CG == 2) mbar_arrive_remote(ready_leader + 8 * s);
                            else mbar_arrive(ready0 + 8 * s);
                        }
                    }
                    if (PASSES == 3) {
                        mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
                        const uint32_t src = tiles + s * stage_bytes + t * 16u, dst = src + off_alo;
                        constexpr uint32_t kIters = (BM * BK * 4) / (32 * kSplitWarps * 16);      // 4
                        float4 v[kIters];
#pragma unroll
                        for (uint32_t i = 0; i < kIters; ++i)
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                                         : "r"(src + i * (32u * kSplitWarps * 16u)));
#pragma unroll
                        for (uint32_t i = 0; i < kIters; ++i) {
                            const float4 h = make_float4(tf32_hi(v[i].x), tf32_hi(v[i].y), tf32_hi(v[i].z), tf32_hi(v[i].w));
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(src + i * (32u * kSplitWarps * 16u)), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w)
                                         : "memory");
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(dst + i * (32u * kSplitWarps * 16u)), "f"(v[i].x - h.x), "f"(v[i].y - h.y),
                                           "f"(v[i].z - h.z), "f"(v[i].w - h.w)
                                         : "memory");
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> tensor core reads
                        named_bar(1, 32 * kSplitWarps);
                        // the arrival travels to the leader CTA: the warps take turns, so that no single thread has one
                        // remote round trip per k-block on its critical path
                        if (t == 32u * (it % kSplitWarps)) {
                            if (CG == 2) mbar_arrive_remote(ready_leader + 8 * s);
                            else mbar_arrive(ready0 + 8 * s);
                        }
                    }
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> swizzled staging tile -> TMA store =====
        const int q = warp & 3;                                     // TMEM lane quarter this warp may read
        const int r_loc = q * 32 + lane;                            // row inside this CTA's 128-row tile
        const uint32_t et = threadIdx.x - 64;                       // 0..127
        const uint32_t tempty_leader = CG == 2 ? map_to_rank(tempty0, 0) : tempty0;
        const uint32_t row_off = (uint32_t)r_loc * 128u, sw = (uint32_t)(r_loc & 7);
        uint32_t tile_it = 0, chunk_it = 0;
        {
            int last_mb = -1, m0 = 0;
            float ep_inv = 0.f;
            const uint32_t* ep_words = nullptr;
            for (int q = 0; sched.get(q, mb, nb); ++q, ++tile_it) {
                if (ep.bias != nullptr) {
                    // the tile's bias -> shared memory while the MMAs of the tile are still running (every epilogue warp is
                    // past the last chunk barrier of the previous tile, i.e. past its last read of s_bias)
                    for (int c = (int)et; c < n_tile; c += 32 * kEpiWarps) {
                        const int col = nb * n_tile + c;
                        s_bias[c] = col < N ? __ldg(ep.bias + col) : 0.f;
                    }
                    named_bar(2, 32 * kEpiWarps);
                }
                if (mb != last_mb) {
                    last_mb = mb;
                    m0 = (mb * CG + (int)rank) * BM;
                    const int row = m0 + r_loc;
                    ep_words = nullptr;
                    if (ep.act != nullptr && row < M) {
                        const int bb = row / ep.T, tt = row - bb * ep.T;
                        ep_inv = __frcp_rn(ep.denom[row]);
                        ep_words = ep.act + ((size_t)bb * ((N + 31) / 32)) * ep.T + tt;
                    }
                }
                const uint32_t a = tile_it & 1u;
                mbar_wait(tfull0 + 8 * a, (tile_it >> 1) & 1);
                tc_fence_after();
                const int n0 = nb * n_tile;
                const int cols = min(n_tile, N - n0);
                for (int c0 = 0; c0 < cols; c0 += 32, ++chunk_it) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + a * kAccCols + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr)
                        : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c0 + 32 >= cols) {              // last read of this accumulator: hand it back to the MMA thread
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) mbar_arrive_remote(tempty_leader + 8 * a);
                            else mbar_arrive(tempty0 + 8 * a);
                        }
                    }
                    if (ep.bias != nullptr) {
                        const float4* b4 = reinterpret_cast<const float4*>(&s_bias[c0]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = b4[j];
                            v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + b.x);
                            v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
                            v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
                            v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
                        }
                    }
                    if (ep_words != nullptr) {          // g = (dx * (bit * scale)) * (1 / denom), as K2 forms it
                        const int cg = n0 + c0, wi = cg >> 5, sh = cg & 31, nw = (N + 31) / 32;
                        unsigned long long bits = wi < nw ? (unsigned long long)ep_words[(size_t)wi * ep.T] : 0ull;
                        if (sh != 0 && wi + 1 < nw) bits |= (unsigned long long)ep_words[(size_t)(wi + 1) * ep.T] << 32;
                        const uint32_t mk = (uint32_t)(bits >> sh);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float f = (float)((mk >> j) & 1u) * ep.scale;
                            v[j] = __float_as_uint(__uint_as_float(v[j]) * f * ep_inv);
                        }
                    }
                    // staging buffer (chunk_it & 1): the store issued two chunks ago must have finished reading it
                    const uint32_t buf = epi_buf + (chunk_it & 1u) * kEpiBuf;
                    if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    named_bar(2, 32 * kEpiWarps);
#pragma unroll
                    for (int j = 0; j < 8; ++j)         // 16-byte chunk j of the row lands at position j ^ (row % 8)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                     ::"r"(buf + row_off + (((uint32_t)j ^ sw) << 4)), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                                       "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                     : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    named_bar(2, 32 * kEpiWarps);
                    if (et == 0) tma_store_2d(&tm_c, buf, n0 + c0, m0);       // rows >= M / columns >= N are clipped
                }
            }
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();        // the peer may still be signalling barriers in this CTA / reading its operands
    if (warp == 1) tmem_dealloc_all<CG>(tmem_base, 2 * kAccCols);
}

// [rows, cols] fp32 row-major -> store boxes of [128 rows x 32 columns], 128-byte swizzle
inline int make_store_map(CUtensorMap* map, float* base, int rows, int cols) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return GPT_ERR_DRIVER;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GPT_OK : GPT_ERR_DRIVER;
}

inline int sm_count() {
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

template <int PASSES, int CG>
int launch(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_b_lo, const CUtensorMap& tm_c, int M,
           int N, int K, int n_tile, int n_tiles, cudaStream_t st, const MaskEpilogue& ep) {
    const size_t stage = (size_t)(PASSES == 3 ? 2 : 1) * (BM * BK * 4 + (size_t)(n_tile / CG) * BK * (PASSES == 0 ? 2 : 4));
    const size_t budget = 227 * 1024 - 1024 /*alignment*/ - 2 * kEpiBuf - 1536 /*static: barriers + the bias tile*/;
    int stages = (int)(budget / stage);
    stages = stages > kMaxStages ? kMaxStages : stages;
    if (stages < 2) return GPT_ERR_UNSUPPORTED;
    const size_t smem = (size_t)stages * stage + 2 * kEpiBuf + 1024;
    if (int a = gpt_smem_opt_in(gemm_persistent_kernel<PASSES, CG>, smem)) return a;
    const int m_blocks = (M + CG * BM - 1) / (CG * BM);
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    int clusters = sm_count() / CG;
    if (CG == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
        // a persistent grid must be ONE wave: as many clusters as can be co-resident (a GPC with an odd SM count hosts
        // one pair fewer), remembered per device
        static int resident[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64) {
            if (resident[dev] == 0) {
                cfg.gridDim = dim3((unsigned)(clusters * CG));
                cfg.attrs = attr;
                cfg.numAttrs = na;
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, gemm_persistent_kernel<PASSES, CG>, &cfg) != cudaSuccess || n <= 0)
                    n = clusters;
                resident[dev] = n;
                if (getenv("GPT_GEMM_DEBUG")) fprintf(stderr, "gemm_persist<%d,%d>: %d co-resident clusters, %d stages\n", PASSES, CG, n, stages);
            }
            if (resident[dev] < clusters) clusters = resident[dev];
        }
    }
    // few row blocks (or a row count only the device knows): the tiles are dealt round-robin over the clusters (TileSched)
    const int flat = (ep.m_live != nullptr || m_blocks < clusters) ? 1 : 0;
    const long long work = flat ? (long long)m_blocks * n_tiles : m_blocks;
    if (clusters > work) clusters = (int)work;
    cfg.gridDim = dim3((unsigned)(clusters * CG));
    if (g_gpt_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_persistent_kernel<PASSES, CG>, tm_a, tm_b, tm_b_lo, tm_c, M, N, K, n_tile,
                                       n_tiles, stages, ep, flat);
    if (e != cudaSuccess) return (int)e;
    return gpt_launch_status();
}

// Which shapes take this kernel: cta_group 0 = never, 1 = single-CTA tiles (cta_group::1), 2 = CTA pairs (default); rows >=
// min_rows.  Initialised from GPT_GEMM_PERSIST / GPT_GEMM_PERSIST_MIN_ROWS, changed by gpt_gemm_persist_config().
struct Config {
    int cta_group;
    long long min_rows;
};
inline Config& config() {
    static Config c = [] {
        const char* a = getenv("GPT_GEMM_PERSIST");
        const char* b = getenv("GPT_GEMM_PERSIST_MIN_ROWS");
        return Config{a ? atoi(a) : 2, b ? atoll(b) : 65536ll};
    }();
    return c;
}
inline int mode() { return config().cta_group; }
inline long long min_rows() { return config().min_rows; }

// C[M,N] = A[M,K] . B[N,K]^T for large M; returns GPT_ERR_UNSUPPORTED when the shape belongs to the small kernel
inline int run(const float* A, const float* B, const float* b_lo, float* C, int M, int N, int K, cudaStream_t st,
               const MaskEpilogue& ep) {
    const int cg = mode();
    if (cg != 1 && cg != 2) return GPT_ERR_UNSUPPORTED;
    // large M -- or a wide output whose live row count is on the device (the relation-aware layers' shared projection: the
    // static schedule then simply stops at the last live row block, and CTA pairs halve W's L2 -> SMEM traffic)
    const bool wide_live = ep.m_live != nullptr && N >= 2048 && M >= 256;
    if ((M < min_rows() && !wide_live) || N % 4 != 0 || K % 4 != 0 || (reinterpret_cast<uintptr_t>(C) & 15))
        return GPT_ERR_UNSUPPORTED;
    const int n_tiles = (N + 255) / 256;
    const int gran = n_tiles > 1 ? 32 : 16 * cg;                 // a tile that has a neighbour ends on a store-box edge
    int n_tile = ((N + n_tiles - 1) / n_tiles + gran - 1) / gran * gran;
    if (n_tile < 16 * cg) n_tile = 16 * cg;
    if (n_tile > 256) return GPT_ERR_UNSUPPORTED;
    alignas(64) CUtensorMap tm_a, tm_b, tm_b_lo, tm_c;
    int rc = make_map(&tm_a, A, M, K, BM);
    if (rc != GPT_OK) return rc;
    if ((rc = make_map(&tm_b, B, N, K, n_tile / cg)) != GPT_OK) return rc;
    if ((rc = make_map(&tm_b_lo, b_lo ? b_lo : B, N, K, n_tile / cg)) != GPT_OK) return rc;
    if ((rc = make_store_map(&tm_c, C, M, N)) != GPT_OK) return rc;
    if (cg == 2) {
        if (b_lo != nullptr) return launch<3, 2>(tm_a, tm_b, tm_b_lo, tm_c, M, N, K, n_tile, n_tiles, st, ep);
        return launch<1, 2>(tm_a, tm_b, tm_b_lo, tm_c, M, N, K, n_tile, n_tiles, st, ep);
    }
    if (b_lo != nullptr) return launch<3, 1>(tm_a, tm_b, tm_b_lo, tm_c, M, N, K, n_tile, n_tiles, st, ep);
    return launch<1, 1>(tm_a, tm_b, tm_b_lo, tm_c, M, N, K, n_tile, n_tiles, st, ep);
}

// the same for bf16 operands: B16 is bf16 [N, K] (kind::f16, one pass)
inline int run_bf16(const float* A, const void* B16, float* C, int M, int N, int K, cudaStream_t st, const MaskEpilogue& ep) {
    const int cg = mode();
    if (cg != 1 && cg != 2) return GPT_ERR_UNSUPPORTED;
    if (M < min_rows() || N % 4 != 0 || K % 8 != 0 || (reinterpret_cast<uintptr_t>(C) & 15)) return GPT_ERR_UNSUPPORTED;
    const int n_tiles = (N + 255) / 256;
    const int gran = n_tiles > 1 ? 32 : 16 * cg;
    int n_tile = ((N + n_tiles - 1) / n_tiles + gran - 1) / gran * gran;
    if (n_tile < 16 * cg) n_tile = 16 * cg;
    if (n_tile > 256) return GPT_ERR_UNSUPPORTED;
    alignas(64) CUtensorMap tm_a, tm_b, tm_c;
    int rc = make_map(&tm_a, A, M, K, BM);
    if (rc != GPT_OK) return rc;
    if ((rc = make_map_bf16(&tm_b, B16, N, K, n_tile / cg)) != GPT_OK) return rc;
    if ((rc = make_store_map(&tm_c, C, M, N)) != GPT_OK) return rc;
    if (cg == 2) return launch<0, 2>(tm_a, tm_b, tm_b, tm_c, M, N, K, n_tile, n_tiles, st, ep);
    return launch<0, 1>(tm_a, tm_b, tm_b, tm_c, M, N, K, n_tile, n_tiles, st, ep);
}

}  // namespace persist
