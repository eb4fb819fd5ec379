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
//   warps 2..9  splitters: every landed fp32 element v -> hi = round_tf32(v) in place, lo = v - hi next to it; rows
//               whose K1 flag is 0 (no gradient) are written as zeros, so dead rows are never accumulated
//   warp 1      one thread issues, per block, hi.hi + lo.hi + hi.lo as tcgen05.mma.kind::tf32 (M = 128, N <= 256 per
//               instruction, K = 8 rows) into the fp32 accumulator in tensor memory
//   flush       every 64 blocks (1024 rows) and at the end, warps 2..9 read the accumulator (tcgen05.ld) and add it
//               into dW with vector reductions (red.global.add.v4.f32); the partial sums of all CTAs meet in L2.
//               Short accumulation chains keep the tensor core's round-toward-zero accumulation below 1e-5 relative
#include "tcgen05_util.cuh"

namespace {

using namespace tc;

constexpr int WG_ROWS = 16;                 // token rows per ring stage (two UMMA K-steps of 8)
constexpr int WG_BOX_BYTES = WG_ROWS * 128;  // one TMA box: 16 rows x 32 fp32
constexpr int WG_NSLICE = 128;              // dW rows per CTA (UMMA M)
constexpr int WG_ABOXES = WG_NSLICE / 32;
constexpr int WG_MAX_STAGES = 4;
constexpr int WG_SPLIT_WARPS = 8;             // splitter / epilogue warps (warps 2..9)
constexpr int WG_THREADS = 64 + 32 * WG_SPLIT_WARPS;
constexpr int WG_FLUSH = 64;                  // k-blocks (x16 rows) accumulated in tensor memory between two flushes

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                    const unsigned char* __restrict__ flags, float* __restrict__ dW, long long M, int N, int K,
                    int n_slices, long long rows_per_part, int kboxes, int tmem_cols, int STAGES,
                    const int* __restrict__ m_live) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[3 * WG_MAX_STAGES + 2];
    __shared__ uint32_t tmem_base_holder;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = (int)(blockIdx.x % n_slices) * WG_NSLICE;
    const long long part = blockIdx.x / n_slices;
    if (m_live != nullptr) {
        // device-side row count: the row ranges are cut here, over the live rows only (the host cut them over all M rows)
        const long long live = *m_live < M ? *m_live : M;
        const long long parts = gridDim.x / n_slices;
        M = live;
        rows_per_part = ((live + WG_ROWS - 1) / WG_ROWS + parts - 1) / parts * WG_ROWS;
    }
    const long long m_begin = part * rows_per_part;
    const long long m_end = m_begin + rows_per_part < M ? m_begin + rows_per_part : M;
    const int nkb = m_end > m_begin ? (int)((m_end - m_begin + WG_ROWS - 1) / WG_ROWS) : 0;
    // stage: [dY hi: 4 boxes | dY lo: 4 boxes | X hi: kboxes | X lo: kboxes]
    const uint32_t a_bytes = WG_ABOXES * WG_BOX_BYTES, b_bytes = (uint32_t)kboxes * WG_BOX_BYTES;
    const uint32_t stage_bytes = 2u * (a_bytes + b_bytes);
    const uint32_t off_alo = a_bytes, off_b = 2u * a_bytes, off_blo = off_b + b_bytes;
    const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
    const uint32_t full0 = smem_addr(&bars[0]), empty0 = smem_addr(&bars[WG_MAX_STAGES]);
    const uint32_t split0 = smem_addr(&bars[2 * WG_MAX_STAGES]), done = smem_addr(&bars[3 * WG_MAX_STAGES]);
    const uint32_t drained = smem_addr(&bars[3 * WG_MAX_STAGES + 1]);

    if (warp == 0 && lane == 0) {
        prefetch_map(&tm_dy);
        prefetch_map(&tm_x);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
            mbar_init(split0 + 8 * s, 32 * WG_SPLIT_WARPS);
        }
        mbar_init(done, 1);
        mbar_init(drained, 32 * WG_SPLIT_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_addr(&tmem_base_holder), (uint32_t)tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_base_holder;

    if (warp == 0) {
        // ===== TMA producer =====
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
                for (int j = 0; j < kboxes; ++j)
                    tma_load_2d(st + off_b + (uint32_t)j * WG_BOX_BYTES, &tm_x, full0 + 8 * s, 32 * j, row);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const int nchunks = (kboxes + 7) / 8;      // <= 256 accumulator columns per instruction
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const int in_group = kb % WG_FLUSH;     // a fresh accumulator every WG_FLUSH blocks (see the epilogue)
                if (in_group == 0 && kb > 0) {
                    mbar_wait(drained, ((kb / WG_FLUSH) - 1) & 1);
                    fence_after();
                }
                mbar_wait(split0 + 8 * s, (kb / STAGES) & 1);
                fence_after();
                const uint32_t st = tiles + (uint32_t)s * stage_bytes;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {   // hi.hi, lo.hi, hi.lo
                    const uint32_t a_base = st + (pass == 1 ? off_alo : 0u);
                    const uint32_t b_base = st + (pass == 2 ? off_blo : off_b);
                    for (int c = 0; c < nchunks; ++c) {
                        const int nb = kboxes - 8 * c < 8 ? kboxes - 8 * c : 8;
                        const uint32_t idesc = make_idesc_tf32(WG_NSLICE, 32 * nb, true, true);
#pragma unroll
                        for (int k8 = 0; k8 < WG_ROWS / 8; ++k8) {   // 8 rows = two 4-row atoms (SBO 512 B); next K-step: +1024 B
                            const uint64_t a_desc = make_desc(a_base + (uint32_t)k8 * 1024u, WG_BOX_BYTES, 512u, kLayoutSw128Base32);
                            const uint64_t b_desc = make_desc(b_base + (uint32_t)c * 8u * WG_BOX_BYTES + (uint32_t)k8 * 1024u,
                                                              WG_BOX_BYTES, 512u, kLayoutSw128Base32);
                            umma_tf32(tmem_base + (uint32_t)c * 256u, a_desc, b_desc, idesc,
                                      (in_group | pass | k8) != 0 ? 1u : 0u);
                        }
                    }
                }
                umma_commit(empty0 + 8 * s);
                if (in_group == WG_FLUSH - 1 || kb == nkb - 1) umma_commit(done);   // accumulator of this group complete
            }
        }
    } else {
        // ===== splitters (+ the accumulator flushes) =====
        // thread u: 16-byte chunk (u & 127) of the boxes u >> 7, (u >> 7) + 2, ...; the chunk's row is (u & 127) / 8
        const uint32_t u = threadIdx.x - 64, t = u & 127u;
        const int j0 = (int)(u >> 7), nboxes = WG_ABOXES + kboxes;
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;              // warps 2..5: first half of the accumulator columns, 6..9: second
        const int cols = 32 * kboxes, c_begin = half * ((kboxes + 1) / 2) * 32;
        const int c_end = half == 0 ? ((kboxes + 1) / 2) * 32 : cols;
        const int n = n0 + q * 32 + lane;
        float* wrow = dW + (size_t)n * K;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const long long row = m_begin + (long long)kb * WG_ROWS + (t >> 3);
            const bool live = row < m_end && (flags == nullptr || flags[row] != 0);
            mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
            const uint32_t st = tiles + (uint32_t)s * stage_bytes + t * 16u;
            for (int jb = j0; jb < nboxes; jb += 8) {   // four boxes per round: the loads are in flight together
                float4 v[4];
                uint32_t src[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = jb + 2 * i;
                    src[i] = st + (j < WG_ABOXES ? (uint32_t)j * WG_BOX_BYTES
                                                 : off_b + (uint32_t)(j - WG_ABOXES) * WG_BOX_BYTES);
                    if (j < nboxes)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(src[i]));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = jb + 2 * i;
                    if (j >= nboxes) break;
                    const uint32_t dst = src[i] + (j < WG_ABOXES ? off_alo : b_bytes);
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
                mbar_arrive(drained);
            }
        }
    }

    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
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
    const int kboxes = (K + 31) / 32;
    int tmem_cols = 32;
    while (tmem_cols < 32 * kboxes) tmem_cols <<= 1;
    const size_t stage = 2 * (size_t)(WG_ABOXES + kboxes) * WG_BOX_BYTES;
    int stages = (int)((220 * 1024) / stage);
    stages = stages > WG_MAX_STAGES ? WG_MAX_STAGES : stages;
    if (stages < 2) return GPT_ERR_UNSUPPORTED;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_slices = (N + WG_NSLICE - 1) / WG_NSLICE;
    long long m_parts = sms / n_slices;
    if (m_parts < 1) m_parts = 1;
    // When one CTA per SM would leave a quarter of the SMs without one (79 slices of a [D*H = 10 000, K] gradient on 148
    // SMs), run TWO CTAs per SM with a two-stage ring each: twice the row ranges, every SM's tensor core fed by two
    // independent pipelines.  Needs both accumulators in tensor memory (2 x <= 256 columns) and both rings in shared memory.
    if (n_slices * m_parts * 4 < 3LL * sms && tmem_cols <= 256 && 2 * (2 * stage + 1024) + 2048 <= 227 * 1024) {
        stages = 2;
        m_parts = 2LL * sms / n_slices;
    }
    const size_t smem = (size_t)stages * stage + 1024;
    const long long blocks16 = (M + WG_ROWS - 1) / WG_ROWS;
    // every row range ends with a flush of its [128, K] partial tile into dW (vector reductions): a range of fewer than
    // 16 blocks spends more on the flush than on the rows (the TACRED-sized batches: a few hundred blocks in all)
    if (m_parts > blocks16 / 16) m_parts = blocks16 / 16 > 0 ? blocks16 / 16 : 1;
    if (m_parts > blocks16) m_parts = blocks16;
    const long long rows_per_part = ((blocks16 + m_parts - 1) / m_parts) * WG_ROWS;
    m_parts = (M + rows_per_part - 1) / rows_per_part;
    alignas(64) CUtensorMap tm_dy, tm_x;
    int rc = tc::make_map_f32(&tm_dy, dy, M, N, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != GPT_OK) return rc;
    if ((rc = tc::make_map_f32(&tm_x, x, M, K, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != GPT_OK) return rc;
    if (int a = gpt_smem_opt_in(wgrad_tf32x3_kernel, smem)) return a;
    wgrad_tf32x3_kernel<<<(unsigned)(n_slices * m_parts), WG_THREADS, smem, (cudaStream_t)stream>>>(
        tm_dy, tm_x, flags, dw, M, N, K, n_slices, rows_per_part, kboxes, tmem_cols, stages, m_live);
    return gpt_launch_status();
}

extern "C" int gpt_linear_wgrad_tf32x3(const float* dy, const float* x, const uint8_t* flags, float* dw, long long M,
                                       int N, int K, void* stream) {
    return gpt_linear_wgrad_tf32x3_rows(dy, x, flags, dw, M, N, K, nullptr, stream);
}
