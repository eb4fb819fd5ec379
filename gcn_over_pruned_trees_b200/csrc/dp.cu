// K8 -- data-parallel gradient exchange fused with clip + SGD, over NVLink peer memory (no NCCL on the step path).
//
// The reference is single-device; sharding sentences over GPUs adds exactly one exchange per step: the gradient
// mean.  A TACRED-shaped step is ~200 us of kernels, so the exchange has to cost microseconds:
// an NCCL all-reduce of the dense [V, 300] word-embedding gradient (60 MB) would be several steps long.  Here every
// rank owns an exchange region (cudaMalloc + cudaIpc, mapped into every peer) and the step ends with three launches:
//
//   dp_push_kernel    each rank stores, with plain coalesced stores over NVLink, into the region of EVERY rank
//                     (itself included, slot = its rank): its flat dense gradient (1.1 MB), the word id of every
//                     token slot that owns a live embedding row (-1 otherwise), those rows themselves, and
//                     slot_of_word[rank][w] = token slot.  It also clears the rank's own G rows / owner marks, so the
//                     next backward starts clean.  No fence and no flag: an in-kernel system-scope fence per CTA cost
//                     ~5 us and was the largest item of this kernel.
//   dp_reduce_kernel  CTA 0 first raises flags[rank] = step on every peer (st.release.sys) -- this grid starts only
//                     after the push grid has completed, so the pushed data has been performed by then -- and every
//                     CTA waits (acquire) until the flags of all ranks carry this step, then works on LOCAL memory
//                     only: flat gradient = sum over ranks in rank order (all ranks' slices in flight before the first
//                     add); rows: a warp takes 32 / W' token-slot entries at a time with lane = (entry, rank), so one
//                     load fetches the slot every rank holds for each word; the lowest (rank, slot) that has a word
//                     adds the rows of the higher ranks in rank order, every load of a row issued before its first
//                     store.  Every rank performs bit-identical additions, so replicas never drift.  Emits the g^2
//                     partial sums.
//   dp_apply_kernel   K7 on the reduced values (mean = sum / W): clip coefficient, p -= lr * coef * g, resets.
//
// Buffers are double-buffered by step parity: a rank can be at most one step ahead of the slowest one (it cannot pass
// the next reduce), so what it pushes for step s+1 never overwrites what a peer still reads for step s.
// One-shot (every rank receives everything) is the right shape for <= 2.3 MB per rank on NVSwitch: (W-1) x 2.3 MB per
// GPU at 900 GB/s is < 20 us at W = 8, with a single synchronisation.  tools/dp_bench.py times the three kernels with
// W virtual ranks on one GPU (W = 8: 136 us -> 55 us with the structure above).
#include "gpt_common.cuh"

namespace {

constexpr int kDpThreads = 256;
constexpr int kDpWarps = kDpThreads / 32;
constexpr int kDpMaxWorld = 8;
constexpr int kDpRowVec = 3;
constexpr int kDpMaxBlocks = 1024;
constexpr int kDpChunk = 8;           // token-slot entries a warp fetches together (reduce / apply)

struct DpLayout {
    int W, cap_rows, E, V;
    long long n_flat;          // floats, padded to a multiple of 4
    size_t off_step, off_done, off_flags, off_nrows, off_dense, off_ids, off_rows, off_slot, bytes;
};

struct DpPeers {
    unsigned char* base[kDpMaxWorld];
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

DpLayout make_layout(int W, int cap_rows, int E, int V, long long n_flat) {
    DpLayout L{};
    L.W = W; L.cap_rows = cap_rows; L.E = E; L.V = V; L.n_flat = (n_flat + 3) / 4 * 4;
    size_t o = 0;
    L.off_step = o; o += 64;                                   // uint64 step (local use)
    L.off_done = o; o += 64;                                   // uint32 done counters [2]
    L.off_flags = o; o += align_up(sizeof(unsigned long long) * W, 256);
    L.off_nrows = o; o += align_up(sizeof(int) * 2 * W, 256);
    L.off_dense = o; o += align_up(sizeof(float) * 2 * W * (size_t)L.n_flat, 256);
    L.off_ids = o; o += align_up(sizeof(int) * 2 * W * (size_t)cap_rows, 256);
    L.off_rows = o; o += align_up(sizeof(float) * 2 * W * (size_t)cap_rows * E, 256);
    L.off_slot = o; o += align_up(sizeof(int) * 2 * W * (size_t)V, 256);
    L.bytes = o;
    return L;
}

__device__ __forceinline__ unsigned long long* dp_flags(const DpLayout& L, unsigned char* base) {
    return reinterpret_cast<unsigned long long*>(base + L.off_flags);
}
__device__ __forceinline__ int* dp_nrows(const DpLayout& L, unsigned char* base, int par) {
    return reinterpret_cast<int*>(base + L.off_nrows) + par * L.W;
}
__device__ __forceinline__ float* dp_dense(const DpLayout& L, unsigned char* base, int par, int r) {
    return reinterpret_cast<float*>(base + L.off_dense) + (size_t)(par * L.W + r) * L.n_flat;
}
__device__ __forceinline__ int* dp_ids(const DpLayout& L, unsigned char* base, int par, int r) {
    return reinterpret_cast<int*>(base + L.off_ids) + (size_t)(par * L.W + r) * L.cap_rows;
}
__device__ __forceinline__ float* dp_rows(const DpLayout& L, unsigned char* base, int par, int r) {
    return reinterpret_cast<float*>(base + L.off_rows) + (size_t)(par * L.W + r) * L.cap_rows * L.E;
}
__device__ __forceinline__ int* dp_slot(const DpLayout& L, unsigned char* base, int par, int r) {
    return reinterpret_cast<int*>(base + L.off_slot) + (size_t)(par * L.W + r) * L.V;
}

// reduce: a warp serves dp_chunk(W) entries at a time, lane = (entry, rank) with dp_pow2(W) lanes per entry
__host__ __device__ __forceinline__ int dp_pow2(int W) { return W <= 1 ? 1 : W <= 2 ? 2 : W <= 4 ? 4 : 8; }
__host__ __device__ __forceinline__ int dp_chunk(int W) {
    const int c = 32 / dp_pow2(W);
    return c < kDpChunk ? c : kDpChunk;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// NVSwitch multicast: one store to the multicast mapping of the exchange regions lands in the region of EVERY rank (the
// switch replicates it), so a rank's contribution leaves its GPU once instead of W - 1 times.  Plain data movement (no
// in-switch arithmetic): the rank-ordered adds stay local and replicas stay bit-identical.
__device__ __forceinline__ void mc_st_f4(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st_f(float* p, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void mc_st_i(int* p, int v) {
    asm volatile("multimem.st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void dp_init_kernel(const DpLayout L, unsigned char* base) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)gridDim.x * blockDim.x;
    int* slot = reinterpret_cast<int*>(base + L.off_slot);
    for (size_t i = tid; i < (size_t)2 * L.W * L.V; i += n) slot[i] = -1;
    int* ids = reinterpret_cast<int*>(base + L.off_ids);
    for (size_t i = tid; i < (size_t)2 * L.W * L.cap_rows; i += n) ids[i] = -1;
    if (tid < (size_t)L.W) dp_flags(L, base)[tid] = 0ull;
    if (tid < (size_t)2 * L.W) reinterpret_cast<int*>(base + L.off_nrows)[tid] = 0;
    if (tid == 0) {
        *reinterpret_cast<unsigned long long*>(base + L.off_step) = 1ull;
        reinterpret_cast<unsigned*>(base + L.off_done)[0] = 0u;
        reinterpret_cast<unsigned*>(base + L.off_done)[1] = 0u;
    }
}

// ---- push ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDpThreads)
dp_push_kernel(const DpLayout L, const DpPeers P, unsigned char* __restrict__ mc, int rank,
               const float* __restrict__ flat_g, float* __restrict__ G, int* __restrict__ owner,
               const long long* __restrict__ words, int n_rows, int topn, int dense_blocks) {
    GPT_PDL_ENTER();        // launched with programmatic serialization: the grid is resident before the backward's tail ends
    unsigned char* self = P.base[rank];
    const unsigned long long step = *reinterpret_cast<const unsigned long long*>(self + L.off_step);
    const int par = (int)(step & 1ull);
    const int W = L.W, tid = threadIdx.x;
    if ((int)blockIdx.x < dense_blocks) {
        const long long n4 = L.n_flat >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(flat_g);
        for (long long i = (long long)blockIdx.x * kDpThreads + tid; i < n4; i += (long long)dense_blocks * kDpThreads) {
            const float4 v = g4[i];
            if (mc != nullptr) mc_st_f4(dp_dense(L, mc, par, rank) + 4 * i, v);
            else for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(dp_dense(L, P.base[q], par, rank))[i] = v;
        }
    } else {
        const int row_blocks = gridDim.x - dense_blocks;
        const int warp = tid >> 5, lane = tid & 31;
        const int E = L.E, E4 = E >> 2;
        for (int i = ((int)blockIdx.x - dense_blocks) * kDpWarps + warp; i < n_rows; i += row_blocks * kDpWarps) {
            const long long w = words[i];
            const bool live = w != 0 && w < topn && owner[w] == i;          // warp-uniform
            if (lane == 0) {
                if (mc != nullptr) mc_st_i(dp_ids(L, mc, par, rank) + i, live ? (int)w : -1);
                else for (int q = 0; q < W; ++q) dp_ids(L, P.base[q], par, rank)[i] = live ? (int)w : -1;
            }
            if (!live) continue;
            float* gr = G + (size_t)w * E;
            if ((E & 3) == 0) {
                float4 v[kDpRowVec];
#pragma unroll
                for (int j = 0; j < kDpRowVec; ++j)
                    v[j] = lane + 32 * j < E4 ? reinterpret_cast<float4*>(gr)[lane + 32 * j] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (mc != nullptr) {
                    float* dst = dp_rows(L, mc, par, rank) + (size_t)i * E;
#pragma unroll
                    for (int j = 0; j < kDpRowVec; ++j)
                        if (lane + 32 * j < E4) mc_st_f4(dst + 4 * (lane + 32 * j), v[j]);
                } else {
                    for (int q = 0; q < W; ++q) {
                        float4* dst = reinterpret_cast<float4*>(dp_rows(L, P.base[q], par, rank) + (size_t)i * E);
#pragma unroll
                        for (int j = 0; j < kDpRowVec; ++j)
                            if (lane + 32 * j < E4) dst[lane + 32 * j] = v[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < kDpRowVec; ++j)
                    if (lane + 32 * j < E4) reinterpret_cast<float4*>(gr)[lane + 32 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = lane + 32 * kDpRowVec; c < E4; c += 32) {      // rows wider than 384 floats
                    const float4 u = reinterpret_cast<float4*>(gr)[c];
                    if (mc != nullptr) mc_st_f4(dp_rows(L, mc, par, rank) + (size_t)i * E + 4 * c, u);
                    else for (int q = 0; q < W; ++q)
                        reinterpret_cast<float4*>(dp_rows(L, P.base[q], par, rank) + (size_t)i * E)[c] = u;
                    reinterpret_cast<float4*>(gr)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                for (int c = lane; c < E; c += 32) {
                    const float u = gr[c];
                    if (mc != nullptr) mc_st_f(dp_rows(L, mc, par, rank) + (size_t)i * E + c, u);
                    else for (int q = 0; q < W; ++q) dp_rows(L, P.base[q], par, rank)[(size_t)i * E + c] = u;
                    gr[c] = 0.f;
                }
            }
            if (lane == 0) {
                if (mc != nullptr) mc_st_i(dp_slot(L, mc, par, rank) + w, i);
                else for (int q = 0; q < W; ++q) dp_slot(L, P.base[q], par, rank)[w] = i;
                owner[w] = 0x7fffffff;
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        if (mc != nullptr) mc_st_i(dp_nrows(L, mc, par) + rank, n_rows);
        else for (int q = 0; q < W; ++q) dp_nrows(L, P.base[q], par)[rank] = n_rows;
    }
    // no fence here: the flags are raised by the NEXT kernel in stream order (dp_signal_kernel, or CTA 0 of
    // dp_reduce_kernel), i.e. after this grid has completed and its stores have drained
}

// flags[rank] = step in every region.  Runs in a kernel that FOLLOWS dp_push_kernel in stream order: a grid starts only
// after the previous one has completed and its stores (peer stores included) have been performed, so whoever
// acquires the flag sees the pushed data.
__device__ __forceinline__ void dp_raise_flags(const DpLayout& L, const DpPeers& P, int rank, unsigned long long step) {
    if ((int)threadIdx.x < L.W) st_release_sys(dp_flags(L, P.base[threadIdx.x]) + rank, step);
}

__global__ void dp_signal_kernel(const DpLayout L, const DpPeers P, int rank) {
    dp_raise_flags(L, P, rank, *reinterpret_cast<const unsigned long long*>(P.base[rank] + L.off_step));
}

__device__ __forceinline__ float dp_block_sum(float v, float* s_red) {
    v = warp_sum_f(v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kDpWarps; ++i) t += s_red[i];
    }
    return t;
}

// entry e of the concatenated token-slot lists of all ranks -> (rank, slot); s_off[r] = first entry of rank r
__device__ __forceinline__ void dp_entry(const int* s_off, int W, int e, int* r, int* i) {
    int rr = 0;
    while (rr + 1 < W && e >= s_off[rr + 1]) ++rr;
    *r = rr;
    *i = e - s_off[rr];
}

// ---- reduce (local memory only, after the flags) -------------------------------------------------------------------
__global__ void __launch_bounds__(kDpThreads, 4)
dp_reduce_kernel(const DpLayout L, const DpPeers P, int rank, int signal, float* __restrict__ flat_g,
                 int dense_blocks, float* __restrict__ partials) {
    // griddepcontrol.wait returns once the push grid has COMPLETED and its stores (peer stores included) are performed:
    // the guarantee the flag protocol below relies on, with or without the programmatic-launch attribute
    GPT_PDL_ENTER();
    __shared__ float s_red[kDpWarps];
    __shared__ int s_off[kDpMaxWorld + 1], s_n[kDpMaxWorld];
    unsigned char* self = P.base[rank];
    const unsigned long long step = *reinterpret_cast<const unsigned long long*>(self + L.off_step);
    const int par = (int)(step & 1ull);
    const int W = L.W, tid = threadIdx.x;
    if (signal && blockIdx.x == 0) dp_raise_flags(L, P, rank, step);
    if (tid < W) {      // a peer that died must not leave this GPU spinning for ever: trap after ~10 s
        const unsigned long long* f = dp_flags(L, self) + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < step) {
            __nanosleep(100);
            if (clock64() - t0 > 20000000000ll) __trap();
        }
        s_n[tid] = min(dp_nrows(L, self, par)[tid], L.cap_rows);
    }
    __syncthreads();
    if (tid <= W) {
        int o = 0;
        for (int r = 0; r < tid; ++r) o += s_n[r];
        s_off[tid] = o;
    }
    __syncthreads();
    float s = 0.f;
    if ((int)blockIdx.x < dense_blocks) {
        const long long n4 = L.n_flat >> 2;
        for (long long i = (long long)blockIdx.x * kDpThreads + tid; i < n4; i += (long long)dense_blocks * kDpThreads) {
            float4 t[kDpMaxWorld];                               // every rank's slice in flight before the first add
#pragma unroll
            for (int r = 0; r < kDpMaxWorld; ++r)
                if (r < W) t[r] = reinterpret_cast<const float4*>(dp_dense(L, self, par, r))[i];
            float4 a = t[0];
#pragma unroll
            for (int r = 1; r < kDpMaxWorld; ++r)
                if (r < W) { a.x += t[r].x; a.y += t[r].y; a.z += t[r].z; a.w += t[r].w; }
            reinterpret_cast<float4*>(flat_g)[i] = a;
            s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
    } else {
        // Token-slot entries of all ranks, kDpChunk per warp: the lanes fetch the ids together, then the live ones are
        // served one after the other with every load of a row (and of its contributions) issued before the first store.
        const int row_blocks = gridDim.x - dense_blocks;
        const int warp = tid >> 5, lane = tid & 31;
        const int E = L.E, E4 = E >> 2;
        const bool vec = (E & 3) == 0 && E4 <= 32 * kDpRowVec;
        const int total = s_off[W];
        const unsigned wmask = (1u << W) - 1u;
        // lane = (entry j of the chunk, rank q): one instruction fetches, for every entry, the slot every rank holds
        // for its word
        const int Wp = dp_pow2(W);
        int chunk = 1;                              // entries per warp and pass: as few as the grid allows (see dp_apply_kernel)
        while (chunk < dp_chunk(W) && (long long)chunk * row_blocks * kDpWarps < total) chunk <<= 1;
        const int jl = lane / Wp, ql = lane - jl * Wp;
        const int stride = row_blocks * kDpWarps * chunk;
        for (int base = (((int)blockIdx.x - dense_blocks) * kDpWarps + warp) * chunk; base < total; base += stride) {
            const int e = base + jl;
            int r = 0, i = 0, w = -1;
            if (jl < chunk && e < total) {
                dp_entry(s_off, W, e, &r, &i);
                w = dp_ids(L, self, par, r)[i];
            }
            int sl = -1;
            if (w >= 0 && ql < W) sl = dp_slot(L, self, par, ql)[w];
            const unsigned hasm = __ballot_sync(GPT_FULL_MASK, sl >= 0);
            unsigned livem = __ballot_sync(GPT_FULL_MASK, w >= 0 && ql == 0);
            while (livem) {
                const int j = __ffs(livem) - 1;                       // lane (entry, rank 0)
                livem &= livem - 1u;
                const int rj = __shfl_sync(GPT_FULL_MASK, r, j), ij = __shfl_sync(GPT_FULL_MASK, i, j);
                const unsigned has = (hasm >> j) & wmask;
                if (has & ((1u << rj) - 1u)) {     // a lower rank owns this word: it will pick this row up
                    if (lane == 0) dp_ids(L, self, par, rj)[ij] = -1;
                    continue;
                }
                float* mine = dp_rows(L, self, par, rj) + (size_t)ij * E;
                if (vec) {
                    float4 acc[kDpRowVec];
#pragma unroll
                    for (int k = 0; k < kDpRowVec; ++k)
                        acc[k] = lane + 32 * k < E4 ? reinterpret_cast<const float4*>(mine)[lane + 32 * k]
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                    unsigned oth = has & ~((2u << rj) - 1u);        // higher ranks holding the word, in rank order
                    while (oth) {
                        const int r2 = __ffs(oth) - 1;
                        oth &= oth - 1u;
                        const int s2 = __shfl_sync(GPT_FULL_MASK, sl, (j + r2) & 31);
                        const float4* src = reinterpret_cast<const float4*>(dp_rows(L, self, par, r2) + (size_t)s2 * E);
                        float4 t[kDpRowVec];
#pragma unroll
                        for (int k = 0; k < kDpRowVec; ++k)
                            t[k] = lane + 32 * k < E4 ? src[lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int k = 0; k < kDpRowVec; ++k) {
                            acc[k].x += t[k].x; acc[k].y += t[k].y; acc[k].z += t[k].z; acc[k].w += t[k].w;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kDpRowVec; ++k)
                        if (lane + 32 * k < E4) {
                            reinterpret_cast<float4*>(mine)[lane + 32 * k] = acc[k];
                            s += acc[k].x * acc[k].x + acc[k].y * acc[k].y + acc[k].z * acc[k].z + acc[k].w * acc[k].w;
                        }
                } else {
                    int slots[kDpMaxWorld];
#pragma unroll
                    for (int r2 = 0; r2 < kDpMaxWorld; ++r2) slots[r2] = __shfl_sync(GPT_FULL_MASK, sl, (j + r2) & 31);
                    for (int c = lane; c < E; c += 32) {
                        float a = mine[c];
                        for (int r2 = rj + 1; r2 < W; ++r2)
                            if ((has >> r2) & 1u) a += dp_rows(L, self, par, r2)[(size_t)slots[r2] * E + c];
                        mine[c] = a;
                        s += a * a;
                    }
                }
            }
        }
    }
    const float t = dp_block_sum(s, s_red);
    if (tid == 0) partials[blockIdx.x] = t;
}

// ---- apply ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDpThreads, 6)
dp_apply_kernel(const DpLayout L, unsigned char* __restrict__ self, float* __restrict__ param,
                float* __restrict__ flat_g, float* __restrict__ emb_w, int dense_blocks,
                const float* __restrict__ partials, int n_partials, float max_norm, float lr,
                float* __restrict__ total_norm, unsigned long long* __restrict__ step_counter) {
    GPT_PDL_ENTER();
    __shared__ float s_red[kDpWarps];
    __shared__ float s_coef;
    __shared__ int s_off[kDpMaxWorld + 1];
    const unsigned long long step = *reinterpret_cast<const unsigned long long*>(self + L.off_step);
    const int par = (int)(step & 1ull);
    const int W = L.W, tid = threadIdx.x;
    const float inv_w = 1.0f / (float)W;
    if (tid >= 32 && tid <= 32 + W) {          // (warp 1: off the path of the norm below)
        int o = 0;
        for (int r = 0; r < tid - 32; ++r) o += min(dp_nrows(L, self, par)[r], L.cap_rows);
        s_off[tid - 32] = o;
    }
    {
        float s = 0.f;
        for (int i = tid; i < n_partials; i += kDpThreads) s += partials[i];
        const float t = dp_block_sum(s, s_red);
        if (tid == 0) {
            const float norm = sqrtf(t) * inv_w;
            s_coef = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
            if (blockIdx.x == 0 && total_norm != nullptr) *total_norm = norm;
        }
        __syncthreads();
    }
    const float a = lr * s_coef * inv_w;
    if ((int)blockIdx.x < dense_blocks) {
        const long long n4 = L.n_flat >> 2;
        float4* g4 = reinterpret_cast<float4*>(flat_g);
        float4* p4 = reinterpret_cast<float4*>(param);
        for (long long i = (long long)blockIdx.x * kDpThreads + tid; i < n4; i += (long long)dense_blocks * kDpThreads) {
            const float4 g = g4[i];
            float4 p = p4[i];
            p.x -= a * g.x; p.y -= a * g.y; p.z -= a * g.z; p.w -= a * g.w;
            p4[i] = p;
            g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        const int row_blocks = gridDim.x - dense_blocks;
        const int warp = tid >> 5, lane = tid & 31;
        const int E = L.E, E4 = E >> 2;
        const bool vec = (E & 3) == 0 && E4 <= 32 * kDpRowVec;
        const int total = s_off[W];
        // entries per warp and pass: as few as the grid allows (1 when there are more warps than entries).  A warp serves
        // its live entries one after the other -- each a round trip to HBM for the embedding row -- so a chunk of 8 with
        // 3 live entries is 3 dependent latencies where spreading the entries over idle warps costs one.
        const int warps_total = row_blocks * kDpWarps;
        int chunk = 1;
        while (chunk < kDpChunk && (long long)chunk * warps_total < total) chunk <<= 1;
        const int stride = warps_total * chunk;
        for (int base = (((int)blockIdx.x - dense_blocks) * kDpWarps + warp) * chunk; base < total; base += stride) {
            const int e = base + lane;
            int r = 0, i = 0, w = -1;                            // w < 0: not live, or folded into a lower rank's entry
            if (lane < chunk && e < total) {
                dp_entry(s_off, W, e, &r, &i);
                w = dp_ids(L, self, par, r)[i];
            }
            unsigned livem = __ballot_sync(GPT_FULL_MASK, w >= 0);
            while (livem) {
                const int j = __ffs(livem) - 1;
                livem &= livem - 1u;
                const int wj = __shfl_sync(GPT_FULL_MASK, w, j), rj = __shfl_sync(GPT_FULL_MASK, r, j),
                          ij = __shfl_sync(GPT_FULL_MASK, i, j);
                const float* row = dp_rows(L, self, par, rj) + (size_t)ij * E;
                float* wr = emb_w + (size_t)wj * E;
                if (vec) {
                    float4 g[kDpRowVec], p[kDpRowVec];
#pragma unroll
                    for (int k = 0; k < kDpRowVec; ++k)
                        if (lane + 32 * k < E4) {
                            g[k] = reinterpret_cast<const float4*>(row)[lane + 32 * k];
                            p[k] = reinterpret_cast<const float4*>(wr)[lane + 32 * k];
                        }
#pragma unroll
                    for (int k = 0; k < kDpRowVec; ++k)
                        if (lane + 32 * k < E4) {
                            p[k].x -= a * g[k].x; p[k].y -= a * g[k].y; p[k].z -= a * g[k].z; p[k].w -= a * g[k].w;
                            reinterpret_cast<float4*>(wr)[lane + 32 * k] = p[k];
                        }
                } else {
                    for (int c = lane; c < E; c += 32) wr[c] -= a * row[c];
                }
                if (lane < W && lane >= rj) dp_slot(L, self, par, lane)[wj] = -1;
            }
        }
    }
    // the last CTA to finish opens the next step
    __syncthreads();
    if (tid == 0) {
        unsigned* done = reinterpret_cast<unsigned*>(self + L.off_done) + 1;
        __threadfence();
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {
            *done = 0u;
            *reinterpret_cast<unsigned long long*>(self + L.off_step) = step + 1ull;
            if (step_counter != nullptr) *step_counter += 1ull;
        }
    }
}

// Grid of the reduce / apply kernels: ONE wave.  dp_reduce_kernel keeps 4 CTAs per SM resident (launch bounds), so at most
// 4 x 148 CTAs in total -- a grid sized for the region's capacity (1005 CTAs at W = 8) ran as a full wave plus a tail wave
// of mostly idle CTAs, and both kernels are chains of dependent memory latencies: the second wave doubled them.  Both
// loops are grid-stride, so fewer CTAs only means more entries per warp.
constexpr int kDpWaveBlocks = 4 * 148;

void plan(const DpLayout& L, int* dense_blocks, int* row_blocks) {
    long long db = (L.n_flat / 4 + 2 * kDpThreads - 1) / (2 * kDpThreads);     // two float4 per thread while that fits
    if (db < 1) db = 1;
    if (db > 148) db = 148;
    const int per_block = kDpWarps * dp_chunk(L.W);
    long long rb = ((long long)L.W * L.cap_rows + per_block - 1) / per_block;
    if (rb > kDpWaveBlocks - db) rb = kDpWaveBlocks - db;
    *dense_blocks = (int)db;
    *row_blocks = (int)rb;
}

int check_layout(int W, int cap_rows, int E, int V, long long n_flat) {
    if (W < 1 || W > kDpMaxWorld || cap_rows < 1 || E < 1 || V < 1 || n_flat < 0) return GPT_ERR_BAD_ARG;
    return GPT_OK;
}

}  // namespace

extern "C" long long gpt_dp_region_bytes(int W, int cap_rows, int E, int V, long long n_flat) {
    if (check_layout(W, cap_rows, E, V, n_flat) != GPT_OK) return -1;
    return (long long)make_layout(W, cap_rows, E, V, n_flat).bytes;
}

extern "C" int gpt_dp_partials(int W, int cap_rows, int E, int V, long long n_flat) {
    if (check_layout(W, cap_rows, E, V, n_flat) != GPT_OK) return -1;
    int d, r;
    plan(make_layout(W, cap_rows, E, V, n_flat), &d, &r);
    return d + r;
}

// cudaMalloc'ed (not from a caching allocator: the IPC handle names the whole allocation) + its IPC handle (64 bytes)
extern "C" int gpt_dp_alloc(long long bytes, void** ptr, void* ipc_handle_out) {
    GPT_CHECK_ARG(bytes > 0 && ptr);
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    if (ipc_handle_out != nullptr) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, *ptr);
        if (e != cudaSuccess) return (int)e;
        memcpy(ipc_handle_out, &h, sizeof(h));
    }
    return GPT_OK;
}

extern "C" int gpt_dp_open(const void* ipc_handle, void** ptr) {
    GPT_CHECK_ARG(ipc_handle && ptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? GPT_OK : (int)e;
}

extern "C" int gpt_dp_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? GPT_OK : (int)e;
}

extern "C" int gpt_dp_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? GPT_OK : (int)e;
}

extern "C" int gpt_dp_region_init(void* region, int W, int cap_rows, int E, int V, long long n_flat, void* stream) {
    GPT_CHECK_ARG(region);
    int rc = check_layout(W, cap_rows, E, V, n_flat);
    if (rc != GPT_OK) return rc;
    const DpLayout L = make_layout(W, cap_rows, E, V, n_flat);
    dp_init_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(L, reinterpret_cast<unsigned char*>(region));
    return gpt_launch_status();
}

static int dp_push_impl(void* const* regions, void* multicast, int rank, int W, int cap_rows, int E, int V, long long n_flat,
                        const float* flat_g, float* g_emb, int32_t* owner, const int64_t* words, int n_rows,
                        int topn, void* stream) {
    GPT_CHECK_ARG(regions && flat_g && rank >= 0 && rank < W && n_rows >= 0);
    GPT_CHECK_ARG(n_rows == 0 || (g_emb && owner && words));
    int rc = check_layout(W, cap_rows, E, V, n_flat);
    if (rc != GPT_OK) return rc;
    if (n_rows > cap_rows) return GPT_ERR_UNSUPPORTED;
    GPT_CHECK_ARG((reinterpret_cast<uintptr_t>(flat_g) & 15) == 0);
    const DpLayout L = make_layout(W, cap_rows, E, V, n_flat);
    DpPeers P{};
    for (int q = 0; q < W; ++q) {
        GPT_CHECK_ARG(regions[q]);
        P.base[q] = reinterpret_cast<unsigned char*>(regions[q]);
    }
    int d, r;
    plan(L, &d, &r);
    int rb = (n_rows + kDpWarps - 1) / kDpWarps;
    if (rb > kDpMaxBlocks - 296) rb = kDpMaxBlocks - 296;      // one warp per token slot
    (void)r;
    gpt_launch(dp_push_kernel, dim3(d + rb), dim3(kDpThreads), 0, (cudaStream_t)stream,
               L, P, reinterpret_cast<unsigned char*>(multicast), rank, flat_g, g_emb, owner,
               reinterpret_cast<const long long*>(words), n_rows, topn, d);
    return gpt_launch_status();
}

extern "C" int gpt_dp_push(void* const* regions, int rank, int W, int cap_rows, int E, int V, long long n_flat,
                           const float* flat_g, float* g_emb, int32_t* owner, const int64_t* words, int n_rows,
                           int topn, void* stream) {
    return dp_push_impl(regions, nullptr, rank, W, cap_rows, E, V, n_flat, flat_g, g_emb, owner, words, n_rows, topn, stream);
}

// the same push through an NVSwitch multicast mapping of the W regions (multimem.st): every byte leaves the GPU once
extern "C" int gpt_dp_push_multicast(void* const* regions, void* multicast, int rank, int W, int cap_rows, int E, int V,
                                     long long n_flat, const float* flat_g, float* g_emb, int32_t* owner,
                                     const int64_t* words, int n_rows, int topn, void* stream) {
    GPT_CHECK_ARG(multicast != nullptr && (reinterpret_cast<uintptr_t>(multicast) & 15) == 0);
    return dp_push_impl(regions, multicast, rank, W, cap_rows, E, V, n_flat, flat_g, g_emb, owner, words, n_rows, topn,
                        stream);
}

static int dp_peers(void* const* regions, int W, DpPeers* P) {
    for (int q = 0; q < W; ++q) {
        GPT_CHECK_ARG(regions[q]);
        P->base[q] = reinterpret_cast<unsigned char*>(regions[q]);
    }
    return GPT_OK;
}

extern "C" int gpt_dp_signal(void* const* regions, int rank, int W, int cap_rows, int E, int V, long long n_flat,
                             void* stream) {
    GPT_CHECK_ARG(regions && rank >= 0 && rank < W);
    int rc = check_layout(W, cap_rows, E, V, n_flat);
    if (rc != GPT_OK) return rc;
    DpPeers P{};
    if ((rc = dp_peers(regions, W, &P)) != GPT_OK) return rc;
    dp_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(make_layout(W, cap_rows, E, V, n_flat), P, rank);
    return gpt_launch_status();
}

extern "C" int gpt_dp_reduce(void* const* regions, int rank, int signal, int W, int cap_rows, int E, int V,
                             long long n_flat, float* flat_g, float* partials, void* stream) {
    GPT_CHECK_ARG(regions && rank >= 0 && rank < W && flat_g && partials);
    GPT_CHECK_ARG((reinterpret_cast<uintptr_t>(flat_g) & 15) == 0);
    int rc = check_layout(W, cap_rows, E, V, n_flat);
    if (rc != GPT_OK) return rc;
    const DpLayout L = make_layout(W, cap_rows, E, V, n_flat);
    DpPeers P{};
    if (signal) {
        if ((rc = dp_peers(regions, W, &P)) != GPT_OK) return rc;
    } else {
        GPT_CHECK_ARG(regions[rank]);
        P.base[rank] = reinterpret_cast<unsigned char*>(regions[rank]);
    }
    int d, r;
    plan(L, &d, &r);
    gpt_launch(dp_reduce_kernel, dim3(d + r), dim3(kDpThreads), 0, (cudaStream_t)stream, L, P, rank, signal, flat_g, d,
               partials);
    return gpt_launch_status();
}

extern "C" int gpt_dp_apply(void* region, int W, int cap_rows, int E, int V, long long n_flat, float* param,
                            float* flat_g, float* emb_w, const float* partials, float max_norm, float lr,
                            float* total_norm, uint64_t* step_counter, void* stream) {
    GPT_CHECK_ARG(region && param && flat_g && emb_w && partials);
    GPT_CHECK_ARG((reinterpret_cast<uintptr_t>(flat_g) & 15) == 0 && (reinterpret_cast<uintptr_t>(param) & 15) == 0);
    int rc = check_layout(W, cap_rows, E, V, n_flat);
    if (rc != GPT_OK) return rc;
    const DpLayout L = make_layout(W, cap_rows, E, V, n_flat);
    int d, r;
    plan(L, &d, &r);                            // the reduce kernel's grid: that many partial sums to add
    // apply keeps 6 CTAs per SM resident: its own, wider single wave (more warps = fewer entries per warp)
    long long ra = ((long long)L.W * L.cap_rows + kDpWarps - 1) / kDpWarps;
    if (ra > 6 * 148 - d) ra = 6 * 148 - d;
    gpt_launch(dp_apply_kernel, dim3(d + (int)ra), dim3(kDpThreads), 0, (cudaStream_t)stream,
               L, reinterpret_cast<unsigned char*>(region), param, flat_g, emb_w, d, partials, d + r, max_norm, lr,
               total_norm, reinterpret_cast<unsigned long long*>(step_counter));
    return gpt_launch_status();
}
