// K3 weight gradient on the tensor cores (sm_100a):   dW[N,K] += dY[M,N]^T . X[M,K]      (3xTF32, fp32-grade)
//
// (/root/reference/model/gcn.py:270-271 under autograd: the gradient of the layer's nn.Linear weight; the reduction
// runs over all B*T token rows, 2 M of them at the large synthetic shape, where the FFMA kernel of gemm_simt.cu needs
// 50 ms per layer.)
//
// The reduction index m is the SLOW index of both operands as they lie in HBM (row-major [M,N] and [M,K]), i.e. both
// are "MN-major" for the tensor core.  tcgen05 takes that layout directly (instruction-descriptor bits 15/16): a TMA
// box of [32 fp32 x 16 rows] with SWIZZLE_128B_ATOM_32B is exactly one column of four canonical MN-major atoms of the
// SWIZZLE_128B_BASE32B layout (4 rows of 128 bytes each), so no transposed copy of the 4 GB activations is ever made.
//
//   CTA (n_slice, m_part): dW rows n0..n0+127 (TMEM lanes) x all K columns (TMEM columns, <= 512), reduced over its
//                          range of token rows; grid = n_slices x m_parts <= one CTA per SM, the four n-slices of one
//                          row range are neighbours in launch order so that X is read from HBM once and from L2 thrice
//   warp 0      TMA producer: per 16-row block, 4 boxes of dY (128 columns) + ceil(K/32) boxes of X into a ring
//   warps 2..9  splitters: every landed fp32 element v -> hi = round_tf32(v) in place, lo = v - hi into one of two lo
//               buffers behind the ring (a stage is held ~3 us, from the TMA issue to the end of its MMAs; a lo buffer
//               only from the split to the end of the MMAs -- the shared memory saved makes the ring twice as deep); rows
//               whose K1 flag is 0 (no gradient) are written as zeros, so dead rows are never accumulated
//   warp 1      one thread issues, per block, hi.hi + lo.hi + hi.lo as tcgen05.mma.kind::tf32 (M = 128, N <= 256 per
//               instruction, K = 8 rows) into the fp32 accumulator in tensor memory
//   flush       every 64 blocks (1024 rows) and at the end, warps 2..9 read the accumulator (tcgen05.ld) and add it
//               into dW with vector reductions (red.global.add.v4.f32); the partial sums of all CTAs meet in L2.
//               Short accumulation chains keep the tensor core's round-toward-zero accumulation below 1e-5 relative
#include "tcgen05_util.cuh"
#include <cstdlib>

namespace {

using namespace tc;

constexpr int WG_ROWS = 16;                 // token rows per ring stage (two UMMA K-steps of 8)
constexpr int WG_BOX_BYTES = WG_ROWS * 128;  // one TMA box: 16 rows x 32 fp32
constexpr int WG_NSLICE = 128;              // dW rows per CTA (UMMA M)
constexpr int WG_ABOXES = WG_NSLICE / 32;
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_LO = 2;                      // lo buffers (see the kernel): written just before the MMAs that read them
constexpr int WG_SPLIT_WARPS = 8;             // splitter / epilogue warps (warps 2..9)
constexpr int WG_THREADS = 64 + 32 * WG_SPLIT_WARPS;
constexpr int WG_FLUSH = 64;                  // k-blocks (x16 rows) accumulated in tensor memory between two flushes

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// ---- CTA pair (cta_group::2) helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_tf32_cg(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (CG == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        umma_tf32(tmem_d, a_desc, b_desc, idesc, acc);
    }
}
// arrive, once every MMA issued so far by this thread has completed, on the barrier at this offset in every CTA of the pair
template <int CG>
__device__ __forceinline__ void commit_cg(uint32_t bar) {
    if (CG == 2) {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"((unsigned short)3)
            : "memory");
    } else {
        umma_commit(bar);
    }
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t holder, uint32_t cols) {
    if (CG == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        tmem_alloc(holder, cols);
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t base, uint32_t cols) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
    else tmem_dealloc(base, cols);
}

// CG = 2: a CTA PAIR owns 256 rows of dW (two neighbouring n-slices, one per CTA) over one row range.  Each CTA stages and
// splits its own 128 columns of dY and only HALF of X's columns; tcgen05.mma.cta_group::2 (M = 256) reads both halves, so
// per CTA the shared-memory traffic of X -- landing, splitting into hi / lo, operand reads of the three passes -- is halved.
// That traffic, not the tensor core, bounds the single-CTA kernel (ncu: 53 % tensor-pipe activity at the large shape).
// kboxes is then the number of 32-column boxes of X THIS CTA stages per chunk-half (see the launcher).
template <int CG>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                    const unsigned char* __restrict__ flags, float* __restrict__ dW, long long M, int N, int K,
                    int n_units, long long rows_per_part, int kboxes, int tmem_cols, int STAGES,
                    const int* __restrict__ m_live) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[4 * WG_MAX_STAGES + 2 + WG_LO];
    __shared__ uint32_t tmem_base_holder;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
    const bool leader = rank == 0;
    const int unit = (int)(blockIdx.x / CG);                 // a CTA (CG = 1) or a CTA pair: (n-slice unit, row range)
    const int n0 = ((unit % n_units) * CG + (int)rank) * WG_NSLICE;
    const long long part = unit / n_units;
    if (m_live != nullptr) {
        // device-side row count: the row ranges are cut here, over the live rows only (the host cut them over all M rows)
        const long long live = *m_live < M ? *m_live : M;
        const long long parts = (gridDim.x / CG) / n_units;
        M = live;
        rows_per_part = ((live + WG_ROWS - 1) / WG_ROWS + parts - 1) / parts * WG_ROWS;
    }
    const long long m_begin = part * rows_per_part;
    const long long m_end = m_begin + rows_per_part < M ? m_begin + rows_per_part : M;
    const int nkb = m_end > m_begin ? (int)((m_end - m_begin + WG_ROWS - 1) / WG_ROWS) : 0;
    // X boxes: kboxes in all (even when CG = 2), accumulated in chunks of <= 8 boxes = 256 tensor-memory columns; of every
    // chunk this CTA stages 1/CG: chunk c holds nb_c boxes, this CTA its boxes [rank * nb_c / CG, (rank + 1) * nb_c / CG)
    constexpr int kChunk = 8 / CG;                            // boxes of a full chunk staged by one CTA
    const int kb_mine = kboxes / CG;
    // ring of STAGES stages [dY: 4 boxes | X: kb_mine boxes]: the landed fp32 values, rewritten in place as their hi parts;
    // behind it WG_LO buffers of the same shape for the lo parts.  A stage is held from the TMA issue to the end of its MMAs
    // (~3 us: what the ring's depth has to cover), a lo buffer only from the split to the end of the MMAs -- two are enough,
    // and the shared memory they do not take makes the ring twice as deep
    const uint32_t a_bytes = WG_ABOXES * WG_BOX_BYTES, b_bytes = (uint32_t)kb_mine * WG_BOX_BYTES;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t off_b = a_bytes;
    const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
    const uint32_t lo_tiles = tiles + (uint32_t)STAGES * stage_bytes;
    const uint32_t full0 = smem_addr(&bars[0]), empty0 = smem_addr(&bars[WG_MAX_STAGES]);
    const uint32_t split0 = smem_addr(&bars[2 * WG_MAX_STAGES]), ready0 = smem_addr(&bars[3 * WG_MAX_STAGES]);
    const uint32_t done = smem_addr(&bars[4 * WG_MAX_STAGES]), drained = smem_addr(&bars[4 * WG_MAX_STAGES + 1]);
    const uint32_t lo_empty0 = smem_addr(&bars[4 * WG_MAX_STAGES + 2]);

    if (warp == 0 && lane == 0) {
        prefetch_map(&tm_dy);
        prefetch_map(&tm_x);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
            mbar_init(split0 + 8 * s, 32 * WG_SPLIT_WARPS);
            mbar_init(ready0 + 8 * s, 1);                     // CG = 2, leader: the peer's operands of stage s are split
        }
        for (int l = 0; l < WG_LO; ++l) mbar_init(lo_empty0 + 8 * l, 1);
        mbar_init(done, 1);
        mbar_init(drained, 32 * WG_SPLIT_WARPS * CG);         // CG = 2: the flush warps of both CTAs arrive on the leader's
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_cg<CG>(smem_addr(&tmem_base_holder), (uint32_t)tmem_cols);
    fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();       // the peer's barriers exist before anything arrives on them
    fence_after();
    const uint32_t tmem_base = tmem_base_holder;

    if (warp == 0) {
        // ===== TMA producer: this CTA's 128 columns of dY and its share of X's columns =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(empty0 + 8 * s, ((kb / STAGES) - 1) & 1);
                const uint32_t st = tiles + (uint32_t)s * stage_bytes;
                const int row = (int)(m_begin + (long long)kb * WG_ROWS);
                mbar_expect_tx(full0 + 8 * s, a_bytes + b_bytes);
#pragma unroll
                for (int j = 0; j < WG_ABOXES; ++j)   // columns past N / rows past M arrive as zeros
                    tma_load_2d(st + (uint32_t)j * WG_BOX_BYTES, &tm_dy, full0 + 8 * s, n0 + 32 * j, row);
                for (int jl = 0; jl < kb_mine; ++jl) {
                    const int c = jl / kChunk, j = jl - c * kChunk;
                    const int nb_c = kboxes - 8 * c < 8 ? kboxes - 8 * c : 8;
                    const int box = 8 * c + (int)rank * (nb_c / CG) + j;       // columns past K arrive as zeros
                    tma_load_2d(st + off_b + (uint32_t)jl * WG_BOX_BYTES, &tm_x, full0 + 8 * s, 32 * box, row);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ===== MMA issuer (the leader drives the tensor cores of both SMs) =====
            const int nchunks = (kboxes + 7) / 8;      // <= 256 accumulator columns per instruction
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const int in_group = kb % WG_FLUSH;     // a fresh accumulator every WG_FLUSH blocks (see the epilogue)
                if (in_group == 0 && kb > 0) {
                    mbar_wait(drained, ((kb / WG_FLUSH) - 1) & 1);
                    fence_after();
                }
                mbar_wait(split0 + 8 * s, (kb / STAGES) & 1);
                if (CG == 2) mbar_wait(ready0 + 8 * s, (kb / STAGES) & 1);
                fence_after();
                const uint32_t st = tiles + (uint32_t)s * stage_bytes;
                const uint32_t lo = lo_tiles + (uint32_t)(kb % WG_LO) * stage_bytes;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {   // hi.hi, lo.hi, hi.lo
                    const uint32_t a_base = pass == 1 ? lo : st;
                    const uint32_t b_base = (pass == 2 ? lo : st) + off_b;
                    for (int c = 0; c < nchunks; ++c) {
                        const int nb = kboxes - 8 * c < 8 ? kboxes - 8 * c : 8;
                        const uint32_t idesc = make_idesc_tf32(WG_NSLICE * CG, 32 * nb, true, true);
#pragma unroll
                        for (int k8 = 0; k8 < WG_ROWS / 8; ++k8) {   // 8 rows = two 4-row atoms (SBO 512 B); next K-step: +1024 B
                            const uint64_t a_desc = make_desc(a_base + (uint32_t)k8 * 1024u, WG_BOX_BYTES, 512u, kLayoutSw128Base32);
                            const uint64_t b_desc = make_desc(b_base + (uint32_t)(c * kChunk) * WG_BOX_BYTES + (uint32_t)k8 * 1024u,
                                                              WG_BOX_BYTES, 512u, kLayoutSw128Base32);
                            umma_tf32_cg<CG>(tmem_base + (uint32_t)c * 256u, a_desc, b_desc, idesc,
                                             (in_group | pass | k8) != 0 ? 1u : 0u);
                        }
                    }
                }
                commit_cg<CG>(empty0 + 8 * s);
                commit_cg<CG>(lo_empty0 + 8 * (kb % WG_LO));
                if (in_group == WG_FLUSH - 1 || kb == nkb - 1) commit_cg<CG>(done);   // accumulator of this group complete
            }
        } else if (lane == 0 && CG == 2) {
            // ===== peer CTA: tell the leader when this CTA's operands of a stage are split =====
            const uint32_t ready_leader = map_to_rank(ready0, 0);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(split0 + 8 * s, (kb / STAGES) & 1);
                mbar_arrive_remote(ready_leader + 8 * s);
            }
        }
    } else {
        // ===== splitters (+ the accumulator flushes) =====
        // thread u: 16-byte chunk (u & 127) of the boxes u >> 7, (u >> 7) + 2, ...; the chunk's row is (u & 127) / 8
        const uint32_t u = threadIdx.x - 64, t = u & 127u;
        const int j0 = (int)(u >> 7), nboxes = WG_ABOXES + kb_mine;
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;              // warps 2..5: first half of the accumulator columns, 6..9: second
        const int cols = 32 * kboxes, c_begin = half * ((kboxes + 1) / 2) * 32;
        const int c_end = half == 0 ? ((kboxes + 1) / 2) * 32 : cols;
        const int n = n0 + q * 32 + lane;
        float* wrow = dW + (size_t)n * K;
        const uint32_t drained_leader = CG == 2 ? map_to_rank(drained, 0) : drained;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const long long row = m_begin + (long long)kb * WG_ROWS + (t >> 3);
            const bool live = row < m_end && (flags == nullptr || flags[row] != 0);
            mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
            if (kb >= WG_LO) mbar_wait(lo_empty0 + 8 * (kb % WG_LO), ((kb / WG_LO) - 1) & 1);   // its last readers are done
            const uint32_t st = tiles + (uint32_t)s * stage_bytes + t * 16u;
            const uint32_t lo_off = lo_tiles + (uint32_t)(kb % WG_LO) * stage_bytes - (tiles + (uint32_t)s * stage_bytes);
            for (int jb = j0; jb < nboxes; jb += 8) {   // four boxes per round: the loads are in flight together
                float4 v[4];
                uint32_t src[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = jb + 2 * i;
                    src[i] = st + (uint32_t)j * WG_BOX_BYTES;         // dY boxes, then X boxes: one contiguous stage
                    if (j < nboxes)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(src[i]));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = jb + 2 * i;
                    if (j >= nboxes) break;
                    const uint32_t dst = src[i] + lo_off;
                    const float4 w = live ? v[i] : make_float4(0.f, 0.f, 0.f, 0.f);   // dead rows never accumulate
                    const float4 h = make_float4(tf32_hi(w.x), tf32_hi(w.y), tf32_hi(w.z), tf32_hi(w.w));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(src[i]), "f"(h.x), "f"(h.y), "f"(h.z),
                                 "f"(h.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(w.x - h.x), "f"(w.y - h.y),
                                 "f"(w.z - h.z), "f"(w.w - h.w) : "memory");
                }
            }
            fence_async_smem();
            mbar_arrive(split0 + 8 * s);
            // Flush: the tensor core adds into its fp32 accumulator with round-toward-zero, a bias that grows with the
            // length of the chain (measured: ~7e-9 relative per accumulated row).  Every WG_FLUSH blocks the partial sum
            // moves to dW (fp32 reductions in L2, round-to-nearest) and the accumulator starts again from zero.
            if (kb % WG_FLUSH == WG_FLUSH - 1 || kb == nkb - 1) {
                const int g = kb / WG_FLUSH;
                mbar_wait(done, g & 1);
                fence_after();
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    uint32_t a[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, a);
                    if (n < N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int k = c0 + j;
                            if (k + 3 < K) {
                                red_add_v4(wrow + k, __uint_as_float(a[j]), __uint_as_float(a[j + 1]),
                                           __uint_as_float(a[j + 2]), __uint_as_float(a[j + 3]));
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (k + i < K) atomicAdd(wrow + k + i, __uint_as_float(a[j + i]));
                            }
                        }
                    }
                }
                fence_before();
                if (CG == 2) mbar_arrive_remote(drained_leader);
                else mbar_arrive(drained);
            }
        }
    }

    fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();        // the peer may still be signalling barriers in this CTA / reading its operands
    if (warp == 1) tmem_dealloc_cg<CG>(tmem_base, (uint32_t)tmem_cols);
}

}  // namespace

// dw[N,K] += dy[M,N]^T . x[M,K] over the rows with flags[m] != 0 (flags == NULL: every row).  GPT_ERR_UNSUPPORTED when
// the shape cannot be described to TMA / does not fit tensor memory (K > 512, K or N not a multiple of 4): callers fall
// back to gpt_linear_wgrad_rows_f32.
// m_live: optional device-side row count (gpt_live_rows): only rows m < *m_live are read
extern "C" int gpt_linear_wgrad_tf32x3_rows(const float* dy, const float* x, const uint8_t* flags, float* dw, long long M,
                                            int N, int K, const int32_t* m_live, void* stream) {
    GPT_CHECK_ARG(dy && x && dw && M >= 0 && N >= 1 && K >= 1);
    if (M == 0) return GPT_OK;
    if (K % 4 != 0 || N % 4 != 0 || K > 512 || M > 0x7fffffffLL - 64 || (reinterpret_cast<uintptr_t>(dy) & 15) ||
        (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dw) & 15))
        return GPT_ERR_UNSUPPORTED;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_slices = (N + WG_NSLICE - 1) / WG_NSLICE;
    const long long blocks16 = (M + WG_ROWS - 1) / WG_ROWS;
    alignas(64) CUtensorMap tm_dy, tm_x;
    int rc = tc::make_map_f32(&tm_dy, dy, M, N, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != GPT_OK) return rc;
    if ((rc = tc::make_map_f32(&tm_x, x, M, K, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != GPT_OK) return rc;

    // Long reductions over >= 2 n-slices run on CTA pairs (cta_group::2, see the kernel): X staged and split once per pair.
    // Measured at 2 097 152 rows x 512 columns of dY.  Stand-alone: with the old ring (hi and lo halves in every stage: two
    // stages at K = 512) the pair was 19 % faster (6.29 -> 5.09 ms); with the deep ring + two just-in-time lo buffers both
    // forms run 5.1-5.2 ms at K = 512 (4.31 single, 4.43-4.64 pair at K = 360), 63 % tensor-pipe activity.  INSIDE the large
    // step, where the weight gradients share the SMs with the data-gradient chain, the pair form wins: 38.5 ms (single),
    // 38.0 (pairs at K = 512 only), 37.1 (pairs for both layers) -- half the shared-memory traffic per SM leaves more of
    // the SM to its neighbours.  GPT_WGRAD_PAIR=0 / GPT_WGRAD_PAIR_MIN_K=<K> select the single-CTA form.
    static const bool pair_ok = [] { const char* e = getenv("GPT_WGRAD_PAIR"); return e == nullptr || atoi(e) != 0; }();
    static const int pair_min_k = [] { const char* e = getenv("GPT_WGRAD_PAIR_MIN_K"); return e ? atoi(e) : 0; }();
    if (pair_ok && n_slices >= 2 && M >= 32768 && K >= pair_min_k) {
        const int kboxes = ((K + 31) / 32 + 1) / 2 * 2;          // even: every chunk is halved between the two CTAs
        int tmem_cols = 32;
        while (tmem_cols < 32 * kboxes) tmem_cols <<= 1;
        const size_t stage = (size_t)(WG_ABOXES + kboxes / 2) * WG_BOX_BYTES;      // + WG_LO buffers of the same size
        int stages = (int)((220 * 1024) / stage) - WG_LO;
        stages = stages > WG_MAX_STAGES ? WG_MAX_STAGES : stages;
        const size_t smem = (size_t)(stages + WG_LO) * stage + 1024;
        const int n_pairs = (n_slices + 1) / 2;
        long long m_parts = (sms / 2) / n_pairs;
        if (m_parts < 1) m_parts = 1;
        if (m_parts > blocks16 / 16) m_parts = blocks16 / 16 > 0 ? blocks16 / 16 : 1;
        const long long rows_per_part = ((blocks16 + m_parts - 1) / m_parts) * WG_ROWS;
        m_parts = (M + rows_per_part - 1) / rows_per_part;
        if (tmem_cols <= 512 && stages >= 2) {
            if (int a = gpt_smem_opt_in(wgrad_tf32x3_kernel<2>, smem)) return a;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(n_pairs * m_parts * 2));
            cfg.blockDim = dim3(WG_THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = (cudaStream_t)stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            const cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_tf32x3_kernel<2>, tm_dy, tm_x, flags, dw, M, N, K, n_pairs,
                                                     rows_per_part, kboxes, tmem_cols, stages, m_live);
            if (e != cudaSuccess) return (int)e;
            return gpt_launch_status();
        }
    }

    const int kboxes = (K + 31) / 32;
    int tmem_cols = 32;
    while (tmem_cols < 32 * kboxes) tmem_cols <<= 1;
    const size_t stage = (size_t)(WG_ABOXES + kboxes) * WG_BOX_BYTES;              // + WG_LO buffers of the same size
    int stages = (int)((220 * 1024) / stage) - WG_LO;
    stages = stages > WG_MAX_STAGES ? WG_MAX_STAGES : stages;
    if (stages < 2) return GPT_ERR_UNSUPPORTED;
    long long m_parts = sms / n_slices;
    if (m_parts < 1) m_parts = 1;
    // When one CTA per SM would leave a quarter of the SMs without one (79 slices of a [D*H = 10 000, K] gradient on 148
    // SMs), run TWO CTAs per SM with a two-stage ring each: twice the row ranges, every SM's tensor core fed by two
    // independent pipelines.  Needs both accumulators in tensor memory (2 x <= 256 columns) and both rings in shared memory.
    if (n_slices * m_parts * 4 < 3LL * sms && tmem_cols <= 256 && 2 * ((2 + WG_LO) * stage + 1024) + 2048 <= 227 * 1024) {
        stages = (int)((110 * 1024) / stage) - WG_LO;
        stages = stages > WG_MAX_STAGES ? WG_MAX_STAGES : stages;
        m_parts = 2LL * sms / n_slices;
    }
    const size_t smem = (size_t)(stages + WG_LO) * stage + 1024;
    // every row range ends with a flush of its [128, K] partial tile into dW (vector reductions): a range of fewer than
    // 16 blocks spends more on the flush than on the rows (the TACRED-sized batches: a few hundred blocks in all)
    if (m_parts > blocks16 / 16) m_parts = blocks16 / 16 > 0 ? blocks16 / 16 : 1;
    if (m_parts > blocks16) m_parts = blocks16;
    const long long rows_per_part = ((blocks16 + m_parts - 1) / m_parts) * WG_ROWS;
    m_parts = (M + rows_per_part - 1) / rows_per_part;
    if (int a = gpt_smem_opt_in(wgrad_tf32x3_kernel<1>, smem)) return a;
    wgrad_tf32x3_kernel<1><<<(unsigned)(n_slices * m_parts), WG_THREADS, smem, (cudaStream_t)stream>>>(
        tm_dy, tm_x, flags, dw, M, N, K, n_slices, rows_per_part, kboxes, tmem_cols, stages, m_live);
    return gpt_launch_status();
}

extern "C" int gpt_linear_wgrad_tf32x3(const float* dy, const float* x, const uint8_t* flags, float* dw, long long M,
                                       int N, int K, void* stream) {
    return gpt_linear_wgrad_tf32x3_rows(dy, x, flags, dw, M, N, K, nullptr, stream);
}
