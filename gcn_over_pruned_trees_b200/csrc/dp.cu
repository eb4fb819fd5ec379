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
//                     slot_of_word[rank][w] = token slot.  Then a system-scope release of flags[rank] = step on every
//                     peer.  It also clears the rank's own G rows / owner marks, so the next backward starts clean.
//   dp_reduce_kernel  waits (acquire) until the flags of all ranks carry this step, then works on LOCAL memory only:
//                     flat gradient = sum over ranks in rank order; for every word, the lowest (rank, slot) that has
//                     it adds the rows of the higher ranks found through slot_of_word, in rank order.  Every rank
//                     performs bit-identical additions, so replicas never drift.  Emits the g^2 partial sums.
//   dp_apply_kernel   K7 on the reduced values (mean = sum / W): clip coefficient, p -= lr * coef * g, resets.
//
// Buffers are double-buffered by step parity: a rank can be at most one step ahead of the slowest one (it cannot pass
// the next reduce), so what it pushes for step s+1 never overwrites what a peer still reads for step s.
// One-shot (every rank receives everything) is the right shape for <= 2.3 MB per rank on NVSwitch: (W-1) x 2.3 MB per
// GPU at 900 GB/s is < 20 us at W = 8, with a single synchronisation.
#include "gpt_common.cuh"

namespace {

constexpr int kDpThreads = 256;
constexpr int kDpWarps = kDpThreads / 32;
constexpr int kDpMaxWorld = 8;
constexpr int kDpRowVec = 3;
constexpr int kDpMaxBlocks = 1024;

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

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
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
dp_push_kernel(const DpLayout L, const DpPeers P, int rank, const float* __restrict__ flat_g, float* __restrict__ G,
               int* __restrict__ owner, const long long* __restrict__ words, int n_rows, int topn, int dense_blocks) {
    unsigned char* self = P.base[rank];
    const unsigned long long step = *reinterpret_cast<const unsigned long long*>(self + L.off_step);
    const int par = (int)(step & 1ull);
    const int W = L.W, tid = threadIdx.x;
    if ((int)blockIdx.x < dense_blocks) {
        const long long n4 = L.n_flat >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(flat_g);
        for (long long i = (long long)blockIdx.x * kDpThreads + tid; i < n4; i += (long long)dense_blocks * kDpThreads) {
            const float4 v = g4[i];
            for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(dp_dense(L, P.base[q], par, rank))[i] = v;
        }
    } else {
        const int row_blocks = gridDim.x - dense_blocks;
        const int warp = tid >> 5, lane = tid & 31;
        const int E = L.E, E4 = E >> 2;
        for (int i = ((int)blockIdx.x - dense_blocks) * kDpWarps + warp; i < n_rows; i += row_blocks * kDpWarps) {
            const long long w = words[i];
            const bool live = w != 0 && w < topn && owner[w] == i;          // warp-uniform
            if (lane == 0)
                for (int q = 0; q < W; ++q) dp_ids(L, P.base[q], par, rank)[i] = live ? (int)w : -1;
            if (!live) continue;
            float* gr = G + (size_t)w * E;
            if ((E & 3) == 0) {
                float4 v[kDpRowVec];
#pragma unroll
                for (int j = 0; j < kDpRowVec; ++j)
                    v[j] = lane + 32 * j < E4 ? reinterpret_cast<float4*>(gr)[lane + 32 * j] : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int q = 0; q < W; ++q) {
                    float4* dst = reinterpret_cast<float4*>(dp_rows(L, P.base[q], par, rank) + (size_t)i * E);
#pragma unroll
                    for (int j = 0; j < kDpRowVec; ++j)
                        if (lane + 32 * j < E4) dst[lane + 32 * j] = v[j];
                }
#pragma unroll
                for (int j = 0; j < kDpRowVec; ++j)
                    if (lane + 32 * j < E4) reinterpret_cast<float4*>(gr)[lane + 32 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = lane + 32 * kDpRowVec; c < E4; c += 32) {      // rows wider than 384 floats
                    const float4 u = reinterpret_cast<float4*>(gr)[c];
                    for (int q = 0; q < W; ++q)
                        reinterpret_cast<float4*>(dp_rows(L, P.base[q], par, rank) + (size_t)i * E)[c] = u;
                    reinterpret_cast<float4*>(gr)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                for (int c = lane; c < E; c += 32) {
                    const float u = gr[c];
                    for (int q = 0; q < W; ++q) dp_rows(L, P.base[q], par, rank)[(size_t)i * E + c] = u;
                    gr[c] = 0.f;
                }
            }
            if (lane == 0) {
                for (int q = 0; q < W; ++q) dp_slot(L, P.base[q], par, rank)[w] = i;
                owner[w] = 0x7fffffff;
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0)
        for (int q = 0; q < W; ++q) dp_nrows(L, P.base[q], par)[rank] = n_rows;
    // publish: every thread makes its stores visible system-wide, the last CTA to finish raises the flags
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        unsigned* done = reinterpret_cast<unsigned*>(self + L.off_done);
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {
            *done = 0u;
            __threadfence_system();
            for (int q = 0; q < W; ++q) st_release_sys(dp_flags(L, P.base[q]) + rank, step);
        }
    }
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
__global__ void __launch_bounds__(kDpThreads)
dp_reduce_kernel(const DpLayout L, unsigned char* __restrict__ self, float* __restrict__ flat_g, int dense_blocks,
                 float* __restrict__ partials) {
    __shared__ float s_red[kDpWarps];
    __shared__ int s_off[kDpMaxWorld + 1];
    const unsigned long long step = *reinterpret_cast<const unsigned long long*>(self + L.off_step);
    const int par = (int)(step & 1ull);
    const int W = L.W, tid = threadIdx.x;
    if (tid < W) {      // a peer that died must not leave this GPU spinning for ever: trap after ~10 s
        const unsigned long long* f = dp_flags(L, self) + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < step) {
            __nanosleep(200);
            if (clock64() - t0 > 20000000000ll) __trap();
        }
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int r = 0; r < W; ++r) { s_off[r] = o; o += min(dp_nrows(L, self, par)[r], L.cap_rows); }
        s_off[W] = o;
    }
    __syncthreads();
    float s = 0.f;
    if ((int)blockIdx.x < dense_blocks) {
        const long long n4 = L.n_flat >> 2;
        for (long long i = (long long)blockIdx.x * kDpThreads + tid; i < n4; i += (long long)dense_blocks * kDpThreads) {
            float4 a = reinterpret_cast<const float4*>(dp_dense(L, self, par, 0))[i];
            for (int r = 1; r < W; ++r) {
                const float4 b = reinterpret_cast<const float4*>(dp_dense(L, self, par, r))[i];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            reinterpret_cast<float4*>(flat_g)[i] = a;
            s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
    } else {
        const int row_blocks = gridDim.x - dense_blocks;
        const int warp = tid >> 5, lane = tid & 31;
        const int E = L.E;
        const int total = s_off[W];
        for (int e = ((int)blockIdx.x - dense_blocks) * kDpWarps + warp; e < total; e += row_blocks * kDpWarps) {
            int r, i;
            dp_entry(s_off, W, e, &r, &i);
            int* ids = dp_ids(L, self, par, r);
            const int w = ids[i];
            if (w < 0) continue;
            bool first = true;
            for (int r2 = 0; r2 < r; ++r2) first &= dp_slot(L, self, par, r2)[w] < 0;
            if (!first) {                      // a lower rank owns this word: it will pick this row up
                __syncwarp();
                if (lane == 0) ids[i] = -1;
                continue;
            }
            float* mine = dp_rows(L, self, par, r) + (size_t)i * E;
            for (int c = lane; c < E; c += 32) {
                float a = mine[c];
                for (int r2 = r + 1; r2 < W; ++r2) {
                    const int s2 = dp_slot(L, self, par, r2)[w];
                    if (s2 >= 0) a += dp_rows(L, self, par, r2)[(size_t)s2 * E + c];
                }
                mine[c] = a;
                s += a * a;
            }
        }
    }
    const float t = dp_block_sum(s, s_red);
    if (tid == 0) partials[blockIdx.x] = t;
}

// ---- apply ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDpThreads)
dp_apply_kernel(const DpLayout L, unsigned char* __restrict__ self, float* __restrict__ param,
                float* __restrict__ flat_g, float* __restrict__ emb_w, int dense_blocks,
                const float* __restrict__ partials, int n_partials, float max_norm, float lr,
                float* __restrict__ total_norm, unsigned long long* __restrict__ step_counter) {
    __shared__ float s_red[kDpWarps];
    __shared__ float s_coef;
    __shared__ int s_off[kDpMaxWorld + 1];
    const unsigned long long step = *reinterpret_cast<const unsigned long long*>(self + L.off_step);
    const int par = (int)(step & 1ull);
    const int W = L.W, tid = threadIdx.x;
    const float inv_w = 1.0f / (float)W;
    if (tid == 0) {
        int o = 0;
        for (int r = 0; r < W; ++r) { s_off[r] = o; o += min(dp_nrows(L, self, par)[r], L.cap_rows); }
        s_off[W] = o;
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
        const int E = L.E;
        const int total = s_off[W];
        for (int e = ((int)blockIdx.x - dense_blocks) * kDpWarps + warp; e < total; e += row_blocks * kDpWarps) {
            int r, i;
            dp_entry(s_off, W, e, &r, &i);
            const int w = dp_ids(L, self, par, r)[i];
            if (w < 0) continue;                               // not live, or folded into a lower rank's entry
            const float* row = dp_rows(L, self, par, r) + (size_t)i * E;
            float* wr = emb_w + (size_t)w * E;
            for (int c = lane; c < E; c += 32) wr[c] -= a * row[c];
            if (lane < W && lane >= r) dp_slot(L, self, par, lane)[w] = -1;
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

void plan(const DpLayout& L, int* dense_blocks, int* row_blocks) {
    long long db = (L.n_flat / 4 + kDpThreads * 4 - 1) / (kDpThreads * 4);
    if (db < 1) db = 1;
    if (db > 296) db = 296;
    long long rb = ((long long)L.W * L.cap_rows + kDpWarps - 1) / kDpWarps;
    if (rb > kDpMaxBlocks - 296) rb = kDpMaxBlocks - 296;
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

extern "C" int gpt_dp_push(void* const* regions, int rank, int W, int cap_rows, int E, int V, long long n_flat,
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
    if (rb > r) rb = r;
    dp_push_kernel<<<d + rb, kDpThreads, 0, (cudaStream_t)stream>>>(
        L, P, rank, flat_g, g_emb, owner, reinterpret_cast<const long long*>(words), n_rows, topn, d);
    return gpt_launch_status();
}

extern "C" int gpt_dp_reduce(void* region, int W, int cap_rows, int E, int V, long long n_flat, float* flat_g,
                             float* partials, void* stream) {
    GPT_CHECK_ARG(region && flat_g && partials && (reinterpret_cast<uintptr_t>(flat_g) & 15) == 0);
    int rc = check_layout(W, cap_rows, E, V, n_flat);
    if (rc != GPT_OK) return rc;
    const DpLayout L = make_layout(W, cap_rows, E, V, n_flat);
    int d, r;
    plan(L, &d, &r);
    dp_reduce_kernel<<<d + r, kDpThreads, 0, (cudaStream_t)stream>>>(L, reinterpret_cast<unsigned char*>(region),
                                                                      flat_g, d, partials);
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
    plan(L, &d, &r);
    dp_apply_kernel<<<d + r, kDpThreads, 0, (cudaStream_t)stream>>>(
        L, reinterpret_cast<unsigned char*>(region), param, flat_g, emb_w, d, partials, d + r, max_norm, lr,
        total_norm, reinterpret_cast<unsigned long long*>(step_counter));
    return gpt_launch_status();
}
