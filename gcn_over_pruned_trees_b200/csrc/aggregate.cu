// K2 -- GCN neighbourhood aggregation over the pruned-tree CSR, fused with the layer epilogue.
//
// Replaces, per layer, the reference's dense path
//   Ax = adj.bmm(h); AxW = W(Ax) + W(h)        /root/reference/model/gcn.py:269-271
//   AxW / denom ; relu ; gcn_drop              /root/reference/model/gcn.py:390-393
// using linearity of W (SURVEY.md 8c "restated dataflow"): with y = h W^T already projected by the GEMM,
//   out_i = dropout(relu((sum_{j in row i} y_j + y_i + 2 b) / denom_i))
// (row i of the CSR already contains the 84-valued diagonal, so the self term is counted twice and the bias
// twice, exactly as the reference does).
//
// Mapping.  A CTA owns one sentence and walks HS-column slices of its T projected rows (HS = 4*LPR floats).  A slice
// [T x HS] is staged in shared memory with cp.async -- every row is read from HBM exactly once -- next to the
// sentence's CSR (16-bit indices), one packed {start, length, 1/denom} word per row and a permutation of the rows
// sorted by row length.  LPR lanes serve one node: they gather its <= deg+1 rows out of shared memory with 128-bit
// loads and packed FADD2 adds.  A warp works on 32/LPR lane groups x 4 rows in flight; the rows it takes together
// are neighbours in the length-sorted order, so the warp-uniform trip count is close to every row's own length
// (exhausted slots read a zero row: no divergent control flow, little padded work).  When a sentence tile is
// large (few CTAs fit an SM) the CTA is persistent over the sentence's slices and double-buffers them, so the next
// slice streams in while the current one is gathered and written out; otherwise one slice per CTA and the SM
// overlaps many small CTAs.
// HBM traffic = read y once + write out once + CSR (+ 1 bit per element of activation mask for the backward).
//
// Backward (adjacency is symmetric, so A^T = A and the same CSR is reused):
//   g_i  = gout_i * dropscale * [out_i > 0] / denom_i          (formed in shared memory)
//   dy_j = g_j + sum_{i in row j} g_i ,   dbias = 2 * sum_i g_i
// [out_i > 0] comes from the forward's bit mask when given (reads 1/32 of the bytes) or from `out` itself.
#include "tcgen05_util.cuh"   // mbarrier / TMA wrappers (also pulls in gpt_common.cuh)
#include <cstdlib>
#include <unordered_map>

namespace {

constexpr int kLenBuckets = 16;  // rows are bucketed by min(length, 15) for the length-sorted order

struct AggParams {
    const float* y;      // [B*T, H] projected rows (fwd) / gout (bwd)
    const float* aux;    // bwd: out of the forward pass (used when act_in == nullptr)
    const uint32_t* act_in;   // bwd: activation bits written by the forward, see act_index()
    uint32_t* act_out;        // fwd: optional activation bits
    const int* rowptr;   // [B, T+1]
    const int* col;      // [B, cap]
    const float* denom;  // [B*T]
    const unsigned char* flags;  // [B*T]
    const float* bias;   // [H] (fwd)
    float* out;          // [B*T, H]
    float* dbias;        // [H] (bwd, atomically accumulated)
    float* pool_out;     // fwd, optional: [B, 3H] masked max pools of the layer output (K4 fused, see POOL)
    int* pool_arg;       // fwd, optional: [B, 3H] their argmax rows (-1: empty pool)
    const float* pool_g;       // bwd, optional: d loss / d pooled [B, 3H]: the incoming gradient is built from it,
    const int* pool_arg_in;    //      the argmax rows and the activation bits in shared memory (K4's backward fused)
    const float* drop_mask;              // optional explicit, pre-scaled mask [B*T, H]
    const unsigned long long* rng;       // optional {seed, step} on the device
    int B, T, H, cap, use_adj, nbuf;
    int pre_scaled;      // bwd: the input already is g = gout * dropscale * [out > 0] / denom (fused into its producer)
    unsigned subseq, thresh16;           // dropout: keep iff rand16 >= thresh16
    int drop_bits;                       // random bits spent per element: 16, or 1 when p == 0.5 exactly
    float drop_scale;
    // TMA: when tma_rows > 0 the slices of y are brought in by cp.async.bulk.tensor (one elected thread, boxes of
    // [tma_rows x HS] floats, tma_rows divides T) instead of one cp.async per 16 bytes issued by every thread -- in the
    // cp.async version 17 % of all instructions of the forward kernel were address arithmetic for those copies
    int tma_rows;
    // bwd, optional: every stored row of dy is stored a second time at its position in the batch's compact row list
    // (gpt_live_rows: inv[b*T + i] >= 0), so that the weight gradient can read the live rows as one dense block
    const int* inv;
    float* out_c;
};

#ifdef GPT_HOST_EMULATION   // tests/emu: the same accessors on a host array instead of PTX
#include "emu_smem_ops.h"
#else
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_dst), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Four fp32 lanes held as two packed f32x2 registers: the gather adds with FADD2 (2 instructions per 16 bytes).
struct Pack4 {
    unsigned long long lo, hi;
};
__device__ __forceinline__ void add_pk(Pack4& a, const Pack4 b) {
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a.lo) : "l"(b.lo));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a.hi) : "l"(b.hi));
}
__device__ __forceinline__ void mul_pk(Pack4& a, const float s) {   // all four lanes times one scalar
    unsigned long long s2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(s2) : "f"(s));
    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(a.lo) : "l"(s2));
    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(a.hi) : "l"(s2));
}
// The value is what it was, but the compiler no longer knows where it came from: it is kept in registers instead of
// being recomputed at every use.
#define GPT_OPAQUE(ptr) asm volatile("" : "+l"(ptr))
__device__ __forceinline__ float4 unpack(const Pack4 a) {
    float4 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(a.lo));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.z), "=f"(v.w) : "l"(a.hi));
    return v;
}

// Shared-memory accessors on 32-bit shared-window addresses: keeps all address arithmetic in 32 bits and stops
// the compiler from re-deriving the carve-up inside the hot loops.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ Pack4 lds_pk(uint32_t a) {
    Pack4 v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
// 128-bit store to global memory through a pointer whose address space the compiler cannot see (GPT_OPAQUE)
__device__ __forceinline__ void stg128(void* gptr, const float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(gptr)), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

// ---- shared-memory carve-up (host and device agree through Layout) -------------------------------------------
//   tile [nbuf][T+1][HS]  staged rows; row T is all zeros (target of padded gather slots and of rows past T)
//   red  [GROUPS][HS]     backward only: per-lane-group column sums for dbias
//   bits [nbuf or 1][..]  backward: the slice's activation words (when a bit mask is given);
//                         forward: the slice's dropout keep-words (in-kernel Philox dropout)
//   bias [slices*HS]      forward: 2*bias for every column, zero past H
//   meta [T+1]            .x = CSR row start | (row length << 16), .y = bits of 1/denom (0: unobservable row)
//   perm [..]             row ids sorted by row length (longest first), padded with T
//   col  [4T+4]           16-bit column indices (3T used; the tail is slack so that padded slots read in bounds)
//   pool [NT/32][3][HS]x2 forward with fused pooling: per-warp partial (max, argmax) of the three pools
//   actb [T+1][8]         forward, when the activation mask is written: one byte per (row, lane of the row's group)
//                         holding that lane's 4 activation bits; packed into the row's 32-bit word after the slice
//                         (row T takes the stores of the padded slots, so that the store needs no branch)
struct Layout {
    size_t tile, red, bits, bias, meta, perm, col, actb, pool, total;  // byte offsets
};
__host__ __device__ inline size_t tile_floats(int T, int lpr) { return (size_t)(T + 1) * 4 * lpr; }
__host__ __device__ inline int perm_len(int T, int groups) {
    return ((T + groups * 4 - 1) / (groups * 4)) * groups * 4;
}
__host__ __device__ inline int bits_stride(int T, int lpr) { return ((T * ((4 * lpr + 31) / 32) + 3) / 4) * 4; }
__host__ __device__ inline Layout make_layout(int T, int H, int lpr, int nt, int nbuf, bool fwd, bool bits,
                                              bool actb = false, bool pool = false) {
    const int groups = (nt / 32) * (32 / lpr), hs = 4 * lpr;
    Layout L;
    size_t o = 0;
    L.tile = o; o += (size_t)nbuf * tile_floats(T, lpr) * 4;
    L.red = o;  o += fwd ? 0 : (size_t)groups * hs * 4;
    L.bits = o; o += bits ? (size_t)(fwd ? 1 : nbuf) * bits_stride(T, lpr) * 4 + (fwd ? 16 : 0) : 0;  // fwd: + row T
    L.bias = o; o += fwd ? (size_t)((H + hs - 1) / hs) * hs * 4 : 0;
    L.meta = o; o += (size_t)((T + 2) / 2 * 2) * 8;
    L.perm = o; o += (size_t)perm_len(T, groups) * 2;
    o = (o + 3) / 4 * 4;
    L.col = o;  o += (size_t)(4 * T + 4) * 2;
    o = (o + 7) / 8 * 8;
    L.actb = o; o += actb ? (size_t)(T + 1) * 8 : 0;
    L.pool = o; o += pool ? (size_t)(nt / 32) * 3 * hs * 8 : 0;   // per-warp (value, argmax) partials of the fused pools
    L.total = (o + 15) / 16 * 16;
    return L;
}

// Stage the sentence's CSR (16-bit), the per-row {start, length, 1/denom} words and the length-sorted row order;
// zero the pad rows.  Block-wide (contains barriers).
template <bool FWD, int NT>
__device__ __forceinline__ void stage_meta(const AggParams& p, int b, float* tile0, size_t tile_stride, uint2* meta,
                                           unsigned short* perm, int nperm, unsigned short* colv, int HS) {
    __shared__ int hist[kLenBuckets];
    const int T = p.T;
    const int* rp = p.rowptr + (size_t)b * (T + 1);
    const int nnz = p.use_adj ? min(rp[T], p.cap) : 0;
    const int* cb = p.col + (size_t)b * p.cap;
    if (threadIdx.x < kLenBuckets) hist[threadIdx.x] = 0;
    for (int e = threadIdx.x; e < nnz; e += NT) colv[e] = (unsigned short)cb[e];
    for (int c = threadIdx.x; c < HS * p.nbuf; c += NT)
        tile0[(size_t)(c / HS) * tile_stride + (size_t)T * HS + (c % HS)] = 0.f;
    for (int t = T + threadIdx.x; t < nperm; t += NT) perm[t] = (unsigned short)T;
    if (threadIdx.x == 0) meta[T] = make_uint2(0u, 0u);  // rows past T: empty, unobservable
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += NT) {
        const int st = rp[t], len = p.use_adj ? rp[t + 1] - st : 0;
        float inv = __frcp_rn(p.denom[(size_t)b * T + t]);
        if (FWD && p.flags[(size_t)b * T + t] == 0) inv = 0.f;  // unobservable row (not in tree, not an entity)
        meta[t] = make_uint2((unsigned)st | ((unsigned)len << 16), __float_as_uint(inv));
        atomicAdd(&hist[kLenBuckets - 1 - min(len, kLenBuckets - 1)], 1);  // longest rows first
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < kLenBuckets; ++k) { const int n = hist[k]; hist[k] = run; run += n; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += NT) {
        const int len = (int)(meta[t].x >> 16);
        perm[atomicAdd(&hist[kLenBuckets - 1 - min(len, kLenBuckets - 1)], 1)] = (unsigned short)t;
    }
    // (the caller's first barrier publishes perm / meta / col)
}

// The same through TMA: one thread, T / tma_rows boxes of [tma_rows x HS] floats landing row-major (no swizzle) in the
// tile buffer; columns past H arrive as zeros; completion is counted in bytes on `bar`.
__device__ __forceinline__ void issue_slice_tma(const AggParams& p, const CUtensorMap* tm, uint32_t tile_s, uint32_t bar,
                                                int b, int sl, int HS) {
    tc::fence_async_smem();     // earlier generic-proxy reads / writes of this buffer (ordered by the CTA barrier) first
    tc::mbar_expect_tx(bar, (uint32_t)p.T * (uint32_t)HS * 4u);
    for (int r0 = 0; r0 < p.T; r0 += p.tma_rows)
        tc::tma_load_2d(tile_s + (uint32_t)r0 * (uint32_t)HS * 4u, tm, bar, sl * HS, b * p.T + r0);
}
__device__ __forceinline__ void tma_bars_init(unsigned long long* bars) {
    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_addr(&bars[0]), 1);
        tc::mbar_init(tc::smem_addr(&bars[1]), 1);
        fence_mbar_init();
    }
}

// Issue the asynchronous copy of slice `sl` of src[b] into a tile buffer (caller commits the group).
template <int LPR, int NT, bool ALIGNED>
__device__ __forceinline__ void issue_slice(const float* __restrict__ src_b, uint32_t tile_s, int T, int H, int sl) {
    constexpr int HS = 4 * LPR;
    const int col0 = sl * HS;
    if (ALIGNED) {
        for (int q = threadIdx.x; q < T * LPR; q += NT) {
            const int row = q / LPR, cc = (q % LPR) * 4, c = col0 + cc;
            const uint32_t dst = tile_s + (uint32_t)(row * HS + cc) * 4u;
            if (c < H) cp_async16(dst, src_b + (size_t)row * H + c);
            else sts128(dst, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    } else {
        for (int q = threadIdx.x; q < T * HS; q += NT) {
            const int row = q / HS, cc = q % HS, c = col0 + cc;
            const uint32_t dst = tile_s + (uint32_t)q * 4u;
            if (c < H) cp_async4(dst, src_b + (size_t)row * H + c);
            else sts_f32(dst, 0.f);
        }
    }
}

// acc[q] += the rows listed in CSR row q, four rows in flight per lane group.  The trip count is warp-uniform (max
// row length in the warp; rows taken together have nearly equal lengths), exhausted slots select the zero row:
//   per chain and trip: LDS.U16, ISETP+SEL, IMAD, LDS.128, 2 FADD2.
template <int HS>
__device__ __forceinline__ void gather4(uint32_t tile_lane, uint32_t col_s, int T, const unsigned (&m)[4],
                                        Pack4 (&acc)[4]) {
    int n[4];
    uint32_t c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { n[q] = m[q] >> 16; c[q] = col_s + 2u * (m[q] & 0xffffu); }
    const int trips = __reduce_max_sync(GPT_FULL_MASK, max(max(n[0], n[1]), max(n[2], n[3])));
    for (int k = 0; k < trips; ++k) {
        uint32_t j[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) j[q] = lds_u16(c[q] + 2u * (uint32_t)k);
        Pack4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = lds_pk(tile_lane + ((k < n[q]) ? j[q] : (uint32_t)T) * (HS * 4));
#pragma unroll
        for (int q = 0; q < 4; ++q) add_pk(acc[q], x[q]);
    }
}

template <bool ALIGNED>
__device__ __forceinline__ void store4(float* dst, const float4 v, int c, int H) {
    if (ALIGNED) {
        *reinterpret_cast<float4*>(dst) = v;  // caller checked c < H; H % 4 == 0 keeps the vector inside
    } else {
        if (c < H) dst[0] = v.x;
        if (c + 1 < H) dst[1] = v.y;
        if (c + 2 < H) dst[2] = v.z;
        if (c + 3 < H) dst[3] = v.w;
    }
}

// Activation-bit layout (fwd writes, bwd reads; LPR = 8 kernels only): one 32-bit word per (sentence, 32-column
// slice, row), bit k <-> column 32*slice + k:   word index = (b * ceil(H/32) + slice) * T + row.
__host__ __device__ inline size_t act_index(int b, int T, int H, int sl) {
    return ((size_t)b * ((H + 31) / 32) + sl) * T;
}

enum { DROP_NONE = 0, DROP_PHILOX = 1, DROP_MASK = 2 };

// POOL: the CTA holds every row of its column slice, so the three masked max pools of the layer output
// (/root/reference/model/gcn.py:116-121, K4) fall out of the same pass: each thread keeps (max, argmax) of its 4 columns
// over the rows it serves, lane groups meet by shuffles, warps in shared memory.  Ties keep the smallest row, an empty
// pool yields -1e12 / -1, exactly as pool3_fwd_kernel.  p.out may then be null: the layer output itself is not stored.
// ACT: the activation bit mask is written (LPR == 8 only).  A template parameter, and the store itself free of
// branches: with a run-time test around it the four rows a lane group has in flight became four separate convergence
// regions, each with its own exposed shared-memory latency (1.73 ms against 1.54 ms without the mask at the large shape).
template <int LPR, int NT, bool ALIGNED, int DROP, bool POOL = false, bool ACT = POOL>
__global__ void __launch_bounds__(NT, NT == 512 ? 1 : (POOL ? (NT == 128 ? 4 : 2) : 3)) aggregate_fwd_kernel(const AggParams p, const __grid_constant__ CUtensorMap tm) {
    GPT_PDL_TRIGGER();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int HS = 4 * LPR, RPW = 32 / LPR, GROUPS = (NT / 32) * RPW, WPR = (HS + 31) / 32;
    const int T = p.T, H = p.H;
    const int b = blockIdx.y;
    const int nsl = (H + HS - 1) / HS;
    constexpr bool write_act = ACT && (LPR == 8);
    const Layout L = make_layout(T, H, LPR, NT, p.nbuf, true, DROP == DROP_PHILOX, write_act, POOL);
    float* tile0 = reinterpret_cast<float*>(smem_raw + L.tile);
    uint32_t* keepw = reinterpret_cast<uint32_t*>(smem_raw + L.bits);
    float* bias_sm = reinterpret_cast<float*>(smem_raw + L.bias);
    uint2* meta = reinterpret_cast<uint2*>(smem_raw + L.meta);
    unsigned short* perm = reinterpret_cast<unsigned short*>(smem_raw + L.perm);
    unsigned short* colv = reinterpret_cast<unsigned short*>(smem_raw + L.col);
    const size_t tile_stride = tile_floats(T, LPR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile_s0 = smem_u32(tile0), tile_bytes = (uint32_t)tile_stride * 4u;
    const float* yb = p.y + (size_t)b * T * H;

    // the CSR (K1, joined long before this launch) and the bias are not products of the preceding grid: staged while
    // it drains; y is, so its first slice is requested after the wait
    // Small tiles (NT == 256, a launch-latency-bound step): staged before the wait, so that it overlaps the tail of the
    // preceding grid.  Large tiles (NT == 512, bandwidth-bound): the first slice is requested first and the staging
    // overlaps its flight instead.
    __shared__ __align__(8) unsigned long long tma_bars[2];
    const bool tma = p.tma_rows > 0;
    const uint32_t bar_s0 = tc::smem_addr(&tma_bars[0]);
    if (tma) tma_bars_init(tma_bars);
    auto issue_fwd = [&](int buf, int sl) {
        if (tma) {
            if (threadIdx.x == 0) issue_slice_tma(p, &tm, tile_s0 + (uint32_t)buf * tile_bytes, bar_s0 + 8u * buf, b, sl, HS);
        } else {
            issue_slice<LPR, NT, ALIGNED>(yb, tile_s0 + (uint32_t)buf * tile_bytes, T, H, sl);
            cp_async_commit();
        }
    };
    if (NT == 512) {
        GPT_PDL_WAIT();
        issue_fwd(0, blockIdx.x);
    }
    stage_meta<true, NT>(p, b, tile0, tile_stride, meta, perm, perm_len(T, GROUPS), colv, HS);
    for (int c = threadIdx.x; c < nsl * HS; c += NT) bias_sm[c] = (c < H) ? 2.0f * p.bias[c] : 0.f;
    if (NT != 512) {
        GPT_PDL_WAIT();
        issue_fwd(0, blockIdx.x);
    }

    // ---- per-thread constants ---------------------------------------------------------------------------------
    const int cl = (lane % LPR) * 4, sub = lane / LPR;
    unsigned long long seed = 0, step = 0;
    if (DROP == DROP_PHILOX) { seed = p.rng[0]; step = p.rng[1]; }
    const uint32_t meta_s = smem_u32(meta), col_s = smem_u32(colv), perm_s = smem_u32(perm);
    const uint32_t bias_s = smem_u32(bias_sm), keep_s = smem_u32(keepw), actb_s = smem_u32(smem_raw + L.actb);
    const unsigned thresh = p.thresh16;
    const float dscale = p.drop_scale;

    float pool_v[POOL ? 3 : 1][4];
    int pool_a[POOL ? 3 : 1][4];
    if (POOL) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int v = 0; v < 4; ++v) { pool_v[POOL ? k : 0][v] = -1e12f; pool_a[POOL ? k : 0][v] = -1; }
    }

    int it = 0;
    for (int sl = blockIdx.x; sl < nsl; sl += gridDim.x, ++it) {
        const int nxt = sl + gridDim.x;
        const int col0 = sl * HS;
        if (nxt < nsl) issue_fwd((it + 1) & 1, nxt);  // the next slice streams into the other buffer meanwhile
        if (DROP == DROP_PHILOX) {
            // Dropout keep-words of this slice (one bit per element), drawn while the slice is still in flight.
            // One Philox call = 8 rows x 1 column x 16 bits; the stream depends only on
            // (seed, step, layer, sentence, row, column), never on the tiling.
            if (p.drop_bits == 1) {
                // p = 0.5 (the reference's gcn_dropout default): one random BIT per element.  One Philox call per
                // thread = the keep-words of 4 rows x one 32-column block -- no ballots, 1/64 of the calls.
                for (int idx = threadIdx.x; idx < ((T + 3) >> 2) * WPR; idx += NT) {
                    const int rq = idx / WPR, w = idx - rq * WPR;
                    const Philox4 q = philox4x32((uint32_t)(col0 / 32 + w) | (p.subseq << 20), (uint32_t)rq,
                                                 (uint32_t)b ^ 0x31415926u, (uint32_t)step, (uint32_t)seed,
                                                 (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
                    const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (rq * 4 + r < T) keepw[(rq * 4 + r) * WPR + w] = r4[r];
                }
            } else {
            for (int blk = warp; blk * 8 < T; blk += NT / 32) {
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const int c = col0 + w * 32 + lane;
                    const Philox4 q = philox4x32((uint32_t)c | (p.subseq << 20), (uint32_t)blk, (uint32_t)b,
                                                 (uint32_t)step, (uint32_t)seed,
                                                 (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
                    const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
                    uint32_t mine = 0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const uint32_t bits = (r4[r >> 1] >> ((r & 1) * 16)) & 0xffffu;
                        const uint32_t bal = __ballot_sync(GPT_FULL_MASK, bits >= thresh);
                        if (lane == r) mine = bal;
                    }
                    if (lane < 8 && blk * 8 + lane < T) keepw[(blk * 8 + lane) * WPR + w] = mine;
                }
            }
            }
        }
        if (tma) {
            tc::mbar_wait(bar_s0 + 8u * (uint32_t)(it & 1), (uint32_t)(it >> 1) & 1u);
        } else {
            if (nxt < nsl) cp_async_wait<1>();
            else cp_async_wait<0>();
        }
        __syncthreads();

        const int c_lane = col0 + cl;
        const bool col_ok = c_lane < H;  // lanes past H stay in the loops (warp-wide ops inside) but never store
        const Pack4 bias2 = lds_pk(bias_s + (uint32_t)c_lane * 4u);  // 2*bias, zero past H
        const uint32_t tile_lane = tile_s0 + (uint32_t)(it & 1) * tile_bytes + (uint32_t)cl * 4u;
        // the lane's column of row 0 of this sentence; opaque to the compiler, which otherwise re-derives the 64-bit
        // address from the kernel parameters for every row (6 instructions per row instead of one IMAD.WIDE)
        char* out_lane = reinterpret_cast<char*>(p.out + (size_t)b * T * H + c_lane);
        GPT_OPAQUE(out_lane);
        const uint32_t row_bytes = (uint32_t)H * 4u;
        const float* const mask_lane = (DROP == DROP_MASK) ? p.drop_mask + (size_t)b * T * H + c_lane : nullptr;
        uint32_t* const act_sl = write_act ? p.act_out + act_index(b, T, H, sl) : nullptr;
        const uint32_t actb_lane = actb_s + (uint32_t)(lane % LPR);
        const uint32_t keep_lane = keep_s + (uint32_t)(cl >> 5) * 4u;
        const bool store_out = !POOL || p.out != nullptr;

        // LPR lanes per node, four nodes in flight per lane group, taken in length-sorted order.  The loop is
        // warp-uniform; slots past T map to the zero row / empty meta entry (and to the pad rows of the keep-words
        // and of the activation bytes) and are never stored.  No branch inside: the four rows' epilogues interleave.
        for (int base0 = warp * RPW * 4; base0 < T; base0 += GROUPS * 4) {
            const uint2 pr = lds64(perm_s + (uint32_t)(base0 + sub * 4) * 2u);
            const int rows[4] = {(int)(pr.x & 0xffffu), (int)(pr.x >> 16), (int)(pr.y & 0xffffu), (int)(pr.y >> 16)};
            unsigned mx[4];
            float inv[4];
            Pack4 acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint2 m = lds64(meta_s + (uint32_t)rows[q] * 8u);
                mx[q] = m.x;
                inv[q] = __uint_as_float(m.y);  // 0 for unobservable rows -> output 0
                // kept elements are scaled by 1/(1-p): folded into the row's 1/denom (relu commutes with a positive scale)
                if (DROP == DROP_PHILOX) inv[q] *= dscale;
                // the separate W(h) self term (the CSR row holds the 84-diagonal a second time)
                acc[q] = lds_pk(tile_lane + (uint32_t)rows[q] * (HS * 4));
            }
            gather4<HS>(tile_lane, col_s, T, mx, acc);
            uint32_t kw[4] = {0u, 0u, 0u, 0u};
            if (DROP == DROP_PHILOX) {      // the four rows' keep-words, requested together
#pragma unroll
                for (int q = 0; q < 4; ++q) kw[q] = lds32(keep_lane + (uint32_t)rows[q] * (WPR * 4)) >> (cl & 31);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = rows[q];
                add_pk(acc[q], bias2);
                mul_pk(acc[q], inv[q]);
                const float4 a = unpack(acc[q]);
                float res[4] = {fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)};
                const bool live = (i < T) && col_ok;
                if (DROP == DROP_PHILOX) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) res[v] = ((kw[q] >> v) & 1u) ? res[v] : 0.f;
                }
                if (DROP == DROP_MASK) {
                    if (live) {
                        const float* m = mask_lane + (uint32_t)i * (uint32_t)H;
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            if (c_lane + v < H) res[v] *= m[v];
                    }
                }
                if (live && store_out) {
                    char* const dst = out_lane + (size_t)(uint32_t)i * (size_t)row_bytes;
                    const float4 r4 = make_float4(res[0], res[1], res[2], res[3]);
                    if (ALIGNED) stg128(dst, r4);   // H % 4 == 0 keeps the vector inside the row
                    else store4<false>(reinterpret_cast<float*>(dst), r4, c_lane, H);
                }
                if (POOL && live) {
                    const unsigned f = p.flags[(size_t)b * T + i];       // bit0 in tree, bit1 subject, bit2 object
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (f & (1u << k)) {
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const int kk = POOL ? k : 0;
                                if (res[v] > pool_v[kk][v] || (res[v] == pool_v[kk][v] && i < pool_a[kk][v])) {
                                    pool_v[kk][v] = res[v];
                                    pool_a[kk][v] = i;
                                }
                            }
                        }
                    }
                }
                if (write_act) {
                    // LPR == 8: this lane's 4 activation bits, packed into the row's word after the slice.  Slots past T
                    // write row T of actb; columns past H hold zeros (tile and bias are zero there), so do their bits.
                    const uint32_t nib = (res[0] > 0.f ? 1u : 0u) | (res[1] > 0.f ? 2u : 0u) | (res[2] > 0.f ? 4u : 0u) |
                                         (res[3] > 0.f ? 8u : 0u);
                    sts_u8(actb_lane + (uint32_t)i * 8u, nib);
                }
            }
        }
        __syncthreads();  // everyone is done with this buffer / keep-words before the next iteration refills them
        if (POOL) {      // (one slice per CTA in this mode, see dispatch)
            float* s_pv = reinterpret_cast<float*>(smem_raw + L.pool);
            int* s_pa = reinterpret_cast<int*>(s_pv + (NT / 32) * 3 * HS);
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float val = pool_v[POOL ? k : 0][v];
                    int arg = pool_a[POOL ? k : 0][v];
#pragma unroll
                    for (int o = LPR; o < 32; o <<= 1) {                 // the lane groups of the warp share the columns
                        const float ov = __shfl_xor_sync(GPT_FULL_MASK, val, o);
                        const int oa = __shfl_xor_sync(GPT_FULL_MASK, arg, o);
                        if (ov > val || (ov == val && oa >= 0 && (arg < 0 || oa < arg))) { val = ov; arg = oa; }
                    }
                    if (sub == 0) {
                        s_pv[(warp * 3 + k) * HS + cl + v] = val;
                        s_pa[(warp * 3 + k) * HS + cl + v] = arg;
                    }
                }
            __syncthreads();
            for (int idx = threadIdx.x; idx < 3 * HS; idx += NT) {
                const int k = idx / HS, c = idx - k * HS;
                float val = s_pv[k * HS + c];
                int arg = s_pa[k * HS + c];
                for (int w = 1; w < NT / 32; ++w) {
                    const float ov = s_pv[(w * 3 + k) * HS + c];
                    const int oa = s_pa[(w * 3 + k) * HS + c];
                    if (ov > val || (ov == val && oa >= 0 && (arg < 0 || oa < arg))) { val = ov; arg = oa; }
                }
                if (col0 + c < H) {
                    p.pool_out[(size_t)b * 3 * H + (size_t)k * H + col0 + c] = val;
                    p.pool_arg[(size_t)b * 3 * H + (size_t)k * H + col0 + c] = arg;
                }
            }
        }
        if (write_act) {
            // the row's 8 nibbles -> its 32-bit activation word, stored in row order (coalesced); the barrier after the
            // next slice has landed orders these reads before the next writes of actb
            for (int t = threadIdx.x; t < T; t += NT) {
                const uint2 v = lds64(actb_s + (uint32_t)t * 8u);
                const uint32_t lo = v.x | (v.x >> 4), hi = v.y | (v.y >> 4);
                act_sl[t] = ((lo & 0xffu) | ((lo >> 8) & 0xff00u)) | (((hi & 0xffu) | ((hi >> 8) & 0xff00u)) << 16);
            }
        }
    }
}

template <int LPR, int NT, bool ALIGNED>
__global__ void __launch_bounds__(NT, NT == 512 ? 1 : 3) aggregate_bwd_kernel(const AggParams p, const __grid_constant__ CUtensorMap tm) {
    GPT_PDL_TRIGGER();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int HS = 4 * LPR, RPW = 32 / LPR, GROUPS = (NT / 32) * RPW;
    const int T = p.T, H = p.H;
    const int b = blockIdx.y;
    const int nsl = (H + HS - 1) / HS;
    const bool use_act = (p.act_in != nullptr) && (LPR == 8) && !p.pre_scaled;
    const Layout L = make_layout(T, H, LPR, NT, p.nbuf, false, use_act);
    float* tile0 = reinterpret_cast<float*>(smem_raw + L.tile);
    float* red = reinterpret_cast<float*>(smem_raw + L.red);  // [GROUPS][HS]
    uint32_t* actw = reinterpret_cast<uint32_t*>(smem_raw + L.bits);
    uint2* meta = reinterpret_cast<uint2*>(smem_raw + L.meta);
    unsigned short* perm = reinterpret_cast<unsigned short*>(smem_raw + L.perm);
    unsigned short* colv = reinterpret_cast<unsigned short*>(smem_raw + L.col);
    const size_t tile_stride = tile_floats(T, LPR);
    const int actw_stride = bits_stride(T, LPR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tile_s0 = smem_u32(tile0), tile_bytes = (uint32_t)tile_stride * 4u;
    const uint32_t actw_s0 = smem_u32(actw);
    const size_t base = (size_t)b * T * H;
    const float* gb = p.y + base;
    // 16-byte async copies of the activation words need T % 4 == 0 and an aligned source
    const bool act16 = use_act && (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.act_in) & 15) == 0);

    const bool from_pool = p.pool_g != nullptr;     // the gradient arrives as d(pooled), not as a [B,T,H] tensor
    __shared__ __align__(8) unsigned long long tma_bars[2];
    const bool tma = p.tma_rows > 0 && !from_pool;
    const uint32_t bar_s0 = tc::smem_addr(&tma_bars[0]);
    if (tma) tma_bars_init(tma_bars);
    // the slice's activation words travel in the same cp.async group as the slice itself
    auto issue = [&](int buf, int sl) {
        if (use_act) {
            const uint32_t* src = p.act_in + act_index(b, T, H, sl);
            const uint32_t dst = actw_s0 + (uint32_t)(buf * actw_stride) * 4u;
            if (act16) {
                for (int q = tid; q < T / 4; q += NT) cp_async16(dst + (uint32_t)q * 16u, src + q * 4);
            } else {
                for (int q = tid; q < T; q += NT) cp_async4(dst + (uint32_t)q * 4u, src + q);
            }
        }
        if (!from_pool) {
            if (tma) {
                if (tid == 0) issue_slice_tma(p, &tm, tile_s0 + (uint32_t)buf * tile_bytes, bar_s0 + 8u * buf, b, sl, HS);
            } else {
                issue_slice<LPR, NT, ALIGNED>(gb, tile_s0 + (uint32_t)buf * tile_bytes, T, H, sl);
            }
        }
        cp_async_commit();
    };
    if (NT == 512) {    // large tiles: first slice in flight while the CSR is staged (see the forward)
        GPT_PDL_WAIT();
        issue(0, blockIdx.x);
    }
    stage_meta<false, NT>(p, b, tile0, tile_stride, meta, perm, perm_len(T, GROUPS), colv, HS);
    if (NT != 512) {
        GPT_PDL_WAIT();     // (CSR staged while the preceding grid drains; its product, the incoming gradient, after)
        issue(0, blockIdx.x);
    }

    const int cl = (lane % LPR) * 4, sub = lane / LPR;
    const uint32_t meta_s = smem_u32(meta), col_s = smem_u32(colv), perm_s = smem_u32(perm);
    const float ds = p.drop_scale;

    int it = 0;
    for (int sl = blockIdx.x; sl < nsl; sl += gridDim.x, ++it) {
        const int nxt = sl + gridDim.x;
        if (nxt < nsl) {
            issue((it + 1) & 1, nxt);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        if (tma) tc::mbar_wait(bar_s0 + 8u * (uint32_t)(it & 1), (uint32_t)(it >> 1) & 1u);
        __syncthreads();

        // ---- in place: g = gout * dropscale * [out > 0] / denom ------------------------------------------------
        const int col0 = sl * HS;
        const uint32_t tile_s = tile_s0 + (uint32_t)(it & 1) * tile_bytes;
        const uint32_t actw_s = actw_s0 + (uint32_t)((it & 1) * actw_stride) * 4u;
        if (from_pool) {
            // g = dh * [out > 0] / denom with dh = scatter of d(pooled) to the argmax rows (pool3_bwd_kernel's
            // arithmetic, same order of operations): at most three non-zero entries per column
            for (int q = tid; q < T * LPR; q += NT) sts128(tile_s + (uint32_t)q * 16u, make_float4(0.f, 0.f, 0.f, 0.f));
            __syncthreads();
            if (tid < HS && col0 + tid < H) {
                const size_t o = (size_t)b * 3 * H + col0 + tid;
                float v[3];
                int r[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) { v[k] = p.pool_g[o + (size_t)k * H]; r[k] = p.pool_arg_in[o + (size_t)k * H]; }
                if (r[1] == r[0] && r[1] >= 0) { v[0] += v[1]; r[1] = -1; }
                if (r[2] >= 0) {
                    if (r[2] == r[0]) { v[0] += v[2]; r[2] = -1; }
                    else if (r[2] == r[1]) { v[1] += v[2]; r[2] = -1; }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (r[k] < 0 || r[k] >= T) continue;
                    const uint32_t w = lds32(actw_s + (uint32_t)r[k] * 4u) >> tid;       // LPR == 8: HS == 32
                    const float inv = __uint_as_float(lds64(meta_s + (uint32_t)r[k] * 8u).y);
                    const float g = v[k] * ((float)(w & 1u) * ds) * inv;
                    sts_f32(tile_s + (uint32_t)(r[k] * HS + tid) * 4u, g);
                }
            }
            __syncthreads();
        } else if (!p.pre_scaled && use_act && p.drop_mask == nullptr) {
            // the path the model takes ([out > 0] from the forward's bit mask, in-kernel dropout): a thread keeps its
            // column group and walks rows; the factor dropscale * bit is selected, not converted and multiplied (four
            // I2F per item were a quarter-rate pipe's worth of the whole pass)
            const int cc = (tid % LPR) * 4;
            if (col0 + cc < H) {
                constexpr int RSTEP = NT / LPR;     // rows between a thread's consecutive items
                for (int row0 = tid / LPR; row0 < T; row0 += 4 * RSTEP) {
                    // four items loaded, then computed, then stored: the loads are in flight together (the accessors
                    // are ordered among themselves, so a plain unrolled loop would expose one latency per item)
                    float4 g[4];
                    float inv[4];
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int row = min(row0 + j * RSTEP, T);       // past T: the zero row / the empty meta entry
                        g[j] = lds128(tile_s + (uint32_t)(row * HS + cc) * 4u);
                        inv[j] = __uint_as_float(lds32(meta_s + (uint32_t)row * 8u + 4u));
                        w[j] = lds32(actw_s + (uint32_t)min(row, T - 1) * 4u) >> cc;  // LPR == 8: cc < 32
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // same association order on every path: (gout * m) * inv
                        g[j].x = g[j].x * ((w[j] & 1u) ? ds : 0.f) * inv[j];
                        g[j].y = g[j].y * ((w[j] & 2u) ? ds : 0.f) * inv[j];
                        g[j].z = g[j].z * ((w[j] & 4u) ? ds : 0.f) * inv[j];
                        g[j].w = g[j].w * ((w[j] & 8u) ? ds : 0.f) * inv[j];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (row0 + j * RSTEP < T) sts128(tile_s + (uint32_t)((row0 + j * RSTEP) * HS + cc) * 4u, g[j]);
                }
            }
            __syncthreads();
        } else if (!p.pre_scaled) {
#pragma unroll 2
        for (int q = tid; q < T * LPR; q += NT) {
            const int row = q / LPR, cc = (q % LPR) * 4, c = col0 + cc;
            if (c >= H) continue;  // already zero-filled
            const uint32_t a = tile_s + (uint32_t)(row * HS + cc) * 4u;
            float4 g = lds128(a);
            const float inv = __uint_as_float(lds64(meta_s + (uint32_t)row * 8u).y);
            float f[4];
            if (use_act) {
                const uint32_t w = lds32(actw_s + (uint32_t)row * 4u) >> cc;  // LPR == 8: cc < 32
#pragma unroll
                for (int v = 0; v < 4; ++v) f[v] = (float)((w >> v) & 1u);
            } else {
                const size_t off = base + (size_t)row * H + c;
#pragma unroll
                for (int v = 0; v < 4; ++v) f[v] = (c + v < H && p.aux[off + v] > 0.f) ? 1.f : 0.f;
            }
            if (p.drop_mask != nullptr) {
                const size_t off = base + (size_t)row * H + c;
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (c + v < H) f[v] *= p.drop_mask[off + v];
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) f[v] *= ds;
            }
            g.x = g.x * f[0] * inv;  // same association order on every path: (gout * m) * inv
            g.y = g.y * f[1] * inv;
            g.z = g.z * f[2] * inv;
            g.w = g.w * f[3] * inv;
            sts128(a, g);
        }
        __syncthreads();
        }

        // ---- dy_j = g_j + sum_{i in row j} g_i ; column sums for dbias -------------------------------------------
        const int c_lane = col0 + cl;
        const bool col_ok = c_lane < H;
        const uint32_t tile_lane = tile_s + (uint32_t)cl * 4u;
        char* out_lane = reinterpret_cast<char*>(p.out + base + c_lane);
        GPT_OPAQUE(out_lane);       // (see the forward: one IMAD.WIDE per stored row instead of a re-derived address)
        const uint32_t row_bytes = (uint32_t)H * 4u;
        Pack4 csum;
        csum.lo = 0ull;
        csum.hi = 0ull;
        for (int base0 = warp * RPW * 4; base0 < T; base0 += GROUPS * 4) {  // warp-uniform
            const uint2 pr = lds64(perm_s + (uint32_t)(base0 + sub * 4) * 2u);
            const int rows[4] = {(int)(pr.x & 0xffffu), (int)(pr.x >> 16), (int)(pr.y & 0xffffu), (int)(pr.y >> 16)};
            unsigned mx[4];
            Pack4 acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                mx[q] = lds64(meta_s + (uint32_t)rows[q] * 8u).x;
                acc[q] = lds_pk(tile_lane + (uint32_t)rows[q] * (HS * 4));  // slots past T read the zero row
                add_pk(csum, acc[q]);
            }
            gather4<HS>(tile_lane, col_s, T, mx, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (rows[q] < T && col_ok) {            // (a predicated store, no branch)
                    char* const dst = out_lane + (size_t)(uint32_t)rows[q] * (size_t)row_bytes;
                    if (ALIGNED) stg128(dst, unpack(acc[q]));
                    else store4<false>(reinterpret_cast<float*>(dst), unpack(acc[q]), c_lane, H);
                }
            if (p.out_c != nullptr) {                   // CTA-uniform: the live rows a second time, in compact order
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (rows[q] < T && col_ok) {
                        const int pos = p.inv[(size_t)b * T + rows[q]];
                        if (pos >= 0) store4<ALIGNED>(p.out_c + (size_t)pos * H + c_lane, unpack(acc[q]), c_lane, H);
                    }
            }
        }
        if (p.dbias != nullptr) *reinterpret_cast<float4*>(red + (warp * RPW + sub) * HS + cl) = unpack(csum);
        __syncthreads();  // tile buffer free for the refill; red[] complete
        if (p.dbias != nullptr) {
            // red[] is rewritten only after the two barriers of the next iteration, so this read is safe
            for (int c = tid; c < HS; c += NT) {
                float sum = 0.f;
#pragma unroll 8
                for (int g = 0; g < GROUPS; ++g) sum += red[g * HS + c];
                if (col0 + c < H) atomicAdd(p.dbias + col0 + c, 2.0f * sum);  // the bias enters the layer twice
            }
        }
    }
}

// ---- host-side configuration ---------------------------------------------------------------------------------

struct AggConfig {
    int lpr, nt, nbuf, grid_x;
    size_t smem;
};

// Small sentence tiles: one slice per CTA, many CTAs per SM overlap each other's load / gather / store phases.
// Large tiles (< 4 CTAs would fit an SM): 512-thread persistent CTA, narrowest slice, double buffered.
AggConfig pick_config(const AggParams& p, bool fwd, int force_vec) {
    const int T = p.T, H = p.H, B = p.B;
    const bool act = fwd ? (p.act_out != nullptr) : (p.act_in != nullptr && !p.pre_scaled);  // bit layout: LPR = 8
    const bool philox = fwd && p.rng != nullptr && p.thresh16 > 0 && p.drop_mask == nullptr;
    AggConfig c{};
    auto slices = [&](int lpr) { return (H + 4 * lpr - 1) / (4 * lpr); };
    auto bytes = [&](int lpr, int nt, int nbuf) {
        return make_layout(T, H, lpr, nt, nbuf, fwd, fwd ? philox : (act && lpr == 8), fwd && act && lpr == 8,
                           fwd && p.pool_out != nullptr).total;
    };
    if (force_vec == 1 || ((force_vec == 2 || force_vec == 4) && !act)) {
        c.lpr = 8 * force_vec; c.nt = 256; c.nbuf = 1; c.grid_x = slices(c.lpr);
        c.smem = bytes(c.lpr, 256, 1);
        if (c.smem <= 224 * 1024) return c;
    }
    for (int lpr = act ? 8 : 32; lpr >= 8; lpr >>= 1) {
        const size_t smem = bytes(lpr, 256, 1);
        if (smem > 54 * 1024) continue;                       // want >= 4 CTAs per SM
        if (lpr > 8 && (long)slices(lpr) * B < 2 * 148) continue;  // and enough CTAs to fill the machine
        c.lpr = lpr; c.nt = 256; c.nbuf = 1; c.grid_x = slices(lpr); c.smem = smem;
        if (fwd && p.pool_out != nullptr && lpr == 8) {
            // The fused-pool form holds 24 (max, argmax) registers more: 128 per thread, two 256-thread CTAs per SM.
            // A batch of 50 sentences x 7 slices is 350 CTAs against 296 places: a second, nearly empty wave (11 us
            // against 6 for the plain form in the step's timeline).  CTAs of 128 threads fit four to an SM: one wave.
            const char* e_nt = getenv("GPT_AGG_POOL_NT");          // tuning / test knob: 128 or 256
            const int force_nt = e_nt ? atoi(e_nt) : 0;
            const long ctas = (long)c.grid_x * B;
            if (force_nt == 128 || (force_nt != 256 && ctas > 2 * 148 && ctas <= 4 * 148)) {
                c.nt = 128;
                c.smem = bytes(lpr, 128, 1);
            }
        }
        return c;
    }
    c.lpr = 8; c.nt = 512;
    const int nsl = slices(8);
    int split = (2 * 148 + B - 1) / B;                         // CTAs per sentence needed to fill the machine
    if (const char* e = getenv("GPT_AGG_SPLIT")) split = atoi(e);  // tuning knob (tools/agg_bench.py)
    split = split < 1 ? 1 : (split > nsl ? nsl : split);
    c.grid_x = split;
    c.nbuf = (split < nsl) ? 2 : 1;
    c.smem = bytes(8, 512, c.nbuf);
    if (c.smem > 224 * 1024 && c.nbuf == 2) {                  // cannot double buffer: one slice per CTA
        c.nbuf = 1; c.grid_x = nsl;
        c.smem = bytes(8, 512, 1);
    }
    return c;
}

template <typename K>
int launch_kernel(K kernel, const AggConfig& c, const AggParams& p, const CUtensorMap& tm, cudaStream_t st) {
    if (int a = gpt_smem_opt_in(kernel, c.smem)) return a;    // large dynamic smem, once per (device, kernel)
    gpt_launch(kernel, dim3(c.grid_x, p.B), dim3(c.nt), c.smem, st, p, tm);
    return gpt_launch_status();
}

template <int LPR, int NT, bool ALIGNED>
int launch(bool fwd, const AggConfig& c, const AggParams& p, const CUtensorMap& tm, cudaStream_t st) {
    if (!fwd) return launch_kernel(aggregate_bwd_kernel<LPR, NT, ALIGNED>, c, p, tm, st);
    if (p.pool_out != nullptr) {
        if (LPR == 8 && NT == 256 && p.act_out != nullptr && p.drop_mask == nullptr && !(p.rng != nullptr && p.thresh16 > 0))
            return launch_kernel(aggregate_fwd_kernel<8, 256, ALIGNED, DROP_NONE, true>, c, p, tm, st);
        return GPT_ERR_UNSUPPORTED;
    }
    if (LPR == 8 && p.act_out != nullptr) {     // with the activation bit mask (its layout is that of 8 lanes per row)
        constexpr int L8 = 8;                   // (the other widths name the same kernels: nothing more is instantiated)
        if (p.drop_mask != nullptr)
            return launch_kernel(aggregate_fwd_kernel<L8, NT, ALIGNED, DROP_MASK, false, true>, c, p, tm, st);
        if (p.rng != nullptr && p.thresh16 > 0)
            return launch_kernel(aggregate_fwd_kernel<L8, NT, ALIGNED, DROP_PHILOX, false, true>, c, p, tm, st);
        return launch_kernel(aggregate_fwd_kernel<L8, NT, ALIGNED, DROP_NONE, false, true>, c, p, tm, st);
    }
    if (p.act_out != nullptr) return GPT_ERR_UNSUPPORTED;   // pick_config chooses 8 lanes per row whenever a mask is asked for
    if (p.drop_mask != nullptr) return launch_kernel(aggregate_fwd_kernel<LPR, NT, ALIGNED, DROP_MASK>, c, p, tm, st);
    if (p.rng != nullptr && p.thresh16 > 0)
        return launch_kernel(aggregate_fwd_kernel<LPR, NT, ALIGNED, DROP_PHILOX>, c, p, tm, st);
    return launch_kernel(aggregate_fwd_kernel<LPR, NT, ALIGNED, DROP_NONE>, c, p, tm, st);
}

int dispatch(bool fwd, AggParams& p, int force_vec, cudaStream_t st) {
    if (4 * p.T + 4 > 65535) return GPT_ERR_UNSUPPORTED;  // 16-bit CSR indices / row lengths in shared memory
    const bool aligned = (p.H % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const AggConfig c = pick_config(p, fwd, force_vec);
    if (c.smem > 224 * 1024) return GPT_ERR_UNSUPPORTED;
    p.nbuf = c.nbuf;
    p.tma_rows = 0;
    alignas(64) CUtensorMap tm{};
    static const bool tma_off = [] { const char* e = getenv("GPT_AGG_TMA"); return e != nullptr && atoi(e) == 0; }();
    if (aligned && !tma_off && p.pool_g == nullptr && (long long)p.B * p.T < 0x7fffffffLL) {
        int rows = p.T;
        if (rows > 256) {
            rows = 256;
            while (rows > 1 && p.T % rows != 0) --rows;           // boxes must tile the sentence exactly
        }
        tc::EncodeTiledFn enc = tc::encode_fn();
        if (rows >= 32 || rows == p.T) {
            if (enc != nullptr) {
                const cuuint64_t dims[2] = {(cuuint64_t)p.H, (cuuint64_t)p.B * (cuuint64_t)p.T};
                const cuuint64_t strides[1] = {(cuuint64_t)p.H * sizeof(float)};
                const cuuint32_t box[2] = {(cuuint32_t)(4 * c.lpr), (cuuint32_t)rows};
                const cuuint32_t estr[2] = {1, 1};
                if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.y), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    p.tma_rows = rows;
            }
        }
    }
    if (c.nt == 512) return aligned ? launch<8, 512, true>(fwd, c, p, tm, st) : launch<8, 512, false>(fwd, c, p, tm, st);
    if (c.nt == 128) {      // fused-pool forward only (pick_config)
        if (!fwd || p.pool_out == nullptr || p.act_out == nullptr || c.lpr != 8) return GPT_ERR_UNSUPPORTED;
        return aligned ? launch_kernel(aggregate_fwd_kernel<8, 128, true, DROP_NONE, true>, c, p, tm, st)
                       : launch_kernel(aggregate_fwd_kernel<8, 128, false, DROP_NONE, true>, c, p, tm, st);
    }
    if (aligned) {
        if (c.lpr == 32) return launch<32, 256, true>(fwd, c, p, tm, st);
        if (c.lpr == 16) return launch<16, 256, true>(fwd, c, p, tm, st);
        return launch<8, 256, true>(fwd, c, p, tm, st);
    }
    if (c.lpr == 32) return launch<32, 256, false>(fwd, c, p, tm, st);
    if (c.lpr == 16) return launch<16, 256, false>(fwd, c, p, tm, st);
    return launch<8, 256, false>(fwd, c, p, tm, st);
}

void set_dropout(AggParams& p, float drop_p) {
    unsigned th = (unsigned)(drop_p * 65536.0f + 0.5f);
    p.thresh16 = th > 65535u ? 65535u : th;
    p.drop_scale = (p.thresh16 > 0) ? 65536.0f / (65536.0f - (float)p.thresh16) : 1.0f;
    p.drop_bits = (p.thresh16 == 32768u) ? 1 : 16;
}

}  // namespace

extern "C" int gpt_gcn_aggregate_fwd(const float* y, const int32_t* rowptr, const int32_t* col, const float* denom,
                                     const uint8_t* flags, const float* bias, float* out, uint32_t* act_mask, int B,
                                     int T, int H, int use_adj, float drop_p, const uint64_t* rng_state,
                                     uint32_t subseq, const float* drop_mask, int force_vec, void* stream) {
    GPT_CHECK_ARG(y && rowptr && col && denom && flags && bias && out);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && H < (1 << 20) && drop_p >= 0.f && drop_p < 1.f);
    GPT_CHECK_ARG(!(drop_p > 0.f && rng_state == nullptr));  // dropout needs the device-side {seed, step}
    if (B == 0) return GPT_OK;
    if (B > 65535) return GPT_ERR_UNSUPPORTED;
    AggParams p{};
    p.y = y; p.rowptr = rowptr; p.col = col; p.denom = denom; p.flags = flags; p.bias = bias; p.out = out;
    p.act_out = act_mask;
    p.drop_mask = drop_mask;
    p.rng = (drop_p > 0.f) ? reinterpret_cast<const unsigned long long*>(rng_state) : nullptr;
    p.B = B; p.T = T; p.H = H; p.cap = 3 * T; p.use_adj = use_adj;
    p.subseq = subseq & 0xfffu;
    set_dropout(p, drop_p);
    return dispatch(true, p, force_vec, (cudaStream_t)stream);
}

// K2 forward of the LAST layer fused with K4 (max pooling): writes pooled [B,3H] / argmax [B,3H] (+ the activation
// mask); `out` may be NULL when the layer output itself is not needed.  GPT_ERR_UNSUPPORTED when the sentence tile is
// too large for the one-slice-per-CTA configuration (callers then run gpt_gcn_aggregate_fwd + gpt_pool3_fwd).
extern "C" int gpt_gcn_aggregate_fwd_pool(const float* y, const int32_t* rowptr, const int32_t* col,
                                          const float* denom, const uint8_t* flags, const float* bias, float* out,
                                          uint32_t* act_mask, float* pooled, int32_t* argmax, int B, int T, int H,
                                          int use_adj, void* stream) {
    GPT_CHECK_ARG(y && rowptr && col && denom && flags && bias && act_mask && pooled && argmax);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && H < (1 << 20));
    if (B == 0) return GPT_OK;
    if (B > 65535) return GPT_ERR_UNSUPPORTED;
    AggParams p{};
    p.y = y; p.rowptr = rowptr; p.col = col; p.denom = denom; p.flags = flags; p.bias = bias; p.out = out;
    p.act_out = act_mask; p.pool_out = pooled; p.pool_arg = argmax;
    p.B = B; p.T = T; p.H = H; p.cap = 3 * T; p.use_adj = use_adj;
    set_dropout(p, 0.f);
    return dispatch(true, p, 0, (cudaStream_t)stream);
}

extern "C" int gpt_gcn_aggregate_fwd_pool_supported(int B, int T, int H) {
    if (B < 1 || B > 65535 || T < 1 || H < 1 || 4 * T + 4 > 65535) return 0;
    AggParams p{};
    p.B = B; p.T = T; p.H = H;
    p.act_out = reinterpret_cast<uint32_t*>(1);      // only null-ness is looked at by pick_config
    p.pool_out = reinterpret_cast<float*>(1);
    const AggConfig c = pick_config(p, true, 0);
    return ((c.nt == 256 || c.nt == 128) && c.lpr == 8) ? 1 : 0;
}

extern "C" int gpt_gcn_aggregate_bwd(const float* gout, const float* out, const uint32_t* act_mask,
                                     const int32_t* rowptr, const int32_t* col, const float* denom, float* dy,
                                     float* dbias, int B, int T, int H, int use_adj, float drop_p,
                                     const float* drop_mask, int force_vec, void* stream) {
    GPT_CHECK_ARG(gout && (out || act_mask) && rowptr && col && denom && dy);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && drop_p >= 0.f && drop_p < 1.f);
    if (B == 0) return GPT_OK;
    if (B > 65535) return GPT_ERR_UNSUPPORTED;
    AggParams p{};
    p.y = gout; p.aux = out; p.act_in = act_mask; p.rowptr = rowptr; p.col = col; p.denom = denom; p.out = dy;
    p.dbias = dbias;
    p.drop_mask = drop_mask;
    p.B = B; p.T = T; p.H = H; p.cap = 3 * T; p.use_adj = use_adj;
    // same scale the forward applied to kept elements (dropped ones have out == 0 / a clear activation bit)
    set_dropout(p, drop_mask == nullptr ? drop_p : 0.f);
    return dispatch(false, p, force_vec, (cudaStream_t)stream);
}

// K2 backward of the LAST layer fused with K4's backward (max pooling): the incoming gradient is d(pooled) [B,3H] +
// argmax [B,3H]; g = scatter(d pooled) * [out > 0] / denom is formed in shared memory, never in HBM.
extern "C" int gpt_gcn_aggregate_bwd_pool_c(const float* dpooled, const int32_t* argmax, const uint32_t* act_mask,
                                            const int32_t* rowptr, const int32_t* col, const float* denom, float* dy,
                                            float* dbias, const int32_t* inv, float* dy_compact, int B, int T, int H,
                                            int use_adj, void* stream);

extern "C" int gpt_gcn_aggregate_bwd_pool(const float* dpooled, const int32_t* argmax, const uint32_t* act_mask,
                                          const int32_t* rowptr, const int32_t* col, const float* denom, float* dy,
                                          float* dbias, int B, int T, int H, int use_adj, void* stream) {
    return gpt_gcn_aggregate_bwd_pool_c(dpooled, argmax, act_mask, rowptr, col, denom, dy, dbias, nullptr, nullptr, B, T, H,
                                        use_adj, stream);
}

extern "C" int gpt_gcn_aggregate_bwd_pool_c(const float* dpooled, const int32_t* argmax, const uint32_t* act_mask,
                                            const int32_t* rowptr, const int32_t* col, const float* denom, float* dy,
                                            float* dbias, const int32_t* inv, float* dy_compact, int B, int T, int H,
                                            int use_adj, void* stream) {
    GPT_CHECK_ARG(dpooled && argmax && act_mask && rowptr && col && denom && dy);
    GPT_CHECK_ARG((inv == nullptr) == (dy_compact == nullptr));
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1);
    if (B == 0) return GPT_OK;
    if (B > 65535) return GPT_ERR_UNSUPPORTED;
    AggParams p{};
    p.y = dpooled; p.act_in = act_mask; p.rowptr = rowptr; p.col = col; p.denom = denom; p.out = dy; p.dbias = dbias;
    p.pool_g = dpooled; p.pool_arg_in = argmax;
    p.inv = inv; p.out_c = dy_compact;
    p.B = B; p.T = T; p.H = H; p.cap = 3 * T; p.use_adj = use_adj;
    set_dropout(p, 0.f);
    return dispatch(false, p, 0, (cudaStream_t)stream);
}

// inv / dy_compact (both or neither): also store every live row of dy at inv[row] of dy_compact (see AggParams)
extern "C" int gpt_gcn_aggregate_bwd_pre_c(const float* g, const int32_t* rowptr, const int32_t* col, const float* denom,
                                           float* dy, float* dbias, const int32_t* inv, float* dy_compact, int B, int T,
                                           int H, int use_adj, int force_vec, void* stream) {
    GPT_CHECK_ARG(g && rowptr && col && denom && dy && (inv == nullptr) == (dy_compact == nullptr));
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1);
    if (B == 0) return GPT_OK;
    if (B > 65535) return GPT_ERR_UNSUPPORTED;
    AggParams p{};
    p.y = g; p.rowptr = rowptr; p.col = col; p.denom = denom; p.out = dy; p.dbias = dbias;
    p.inv = inv; p.out_c = dy_compact;
    p.B = B; p.T = T; p.H = H; p.cap = 3 * T; p.use_adj = use_adj;
    p.pre_scaled = 1;
    set_dropout(p, 0.f);
    return dispatch(false, p, force_vec, (cudaStream_t)stream);
}

extern "C" int gpt_gcn_aggregate_bwd_pre(const float* g, const int32_t* rowptr, const int32_t* col, const float* denom,
                                         float* dy, float* dbias, int B, int T, int H, int use_adj, int force_vec,
                                         void* stream) {
    return gpt_gcn_aggregate_bwd_pre_c(g, rowptr, col, denom, dy, dbias, nullptr, nullptr, B, T, H, use_adj, force_vec,
                                       stream);
}
