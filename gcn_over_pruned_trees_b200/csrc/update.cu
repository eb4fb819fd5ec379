// K7 -- global-norm gradient clipping + SGD over one flat parameter buffer and the live word-embedding rows.
//
// Replaces  torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm); optimizer.step(); zero_grad()
// (/root/reference/train.py:224-227 with the shipped --optim sgd, utils/torch_utils.py:93-96) -- about 15 ATen
// launches over 12 small tensors plus a dense [V, 300] embedding gradient -- with two launches:
//   update_sqnorm_kernel : per-CTA partial sums of g^2 over the flat dense gradient and over the word-embedding rows
//                          the batch touched (each live row counted once, by its first token: `owner`)
//   update_apply_kernel  : every CTA re-adds the partials in the same fixed order (deterministic, no atomics, nothing
//                          to zero), coef = min(1, max_norm / (sqrt(sum) + 1e-6)) as clip_grad_norm_, then
//                          p -= lr * coef * g ; g = 0   for the flat buffer and the live rows (G stays all-zero between
//                          steps, owner is reset), so the next step needs no zero_grad pass.
// grad_scale multiplies every gradient first (1/world_size after a summing exchange in data-parallel runs).
// step_counter (optional, the {seed, step} dropout state's step word) is advanced by one: new dropout streams for the
// next step without a separate launch.
#include "gpt_common.cuh"

namespace {

constexpr int kUpdThreads = 256;
constexpr int kUpdWarps = kUpdThreads / 32;
constexpr int kMaxPartials = 1024;
constexpr int kRowVec = 3;           // float4 per lane held in registers: rows up to 384 floats in one round trip

__device__ __forceinline__ float block_sum(float v, float* s_red) {
    v = warp_sum_f(v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kUpdWarps; ++i) t += s_red[i];
    }
    return t;  // valid in thread 0
}

// CTAs [0, dense_blocks): flat gradient; CTAs [dense_blocks, dense_blocks + row_blocks): embedding rows.
__global__ void __launch_bounds__(kUpdThreads)
update_sqnorm_kernel(const float* __restrict__ grad, long long n, const long long* __restrict__ words,
                     const int* __restrict__ owner, const float* __restrict__ g_emb, int n_rows, int E, int topn,
                     int dense_blocks, float* __restrict__ partials) {
    GPT_PDL_ENTER();
    __shared__ float s_red[kUpdWarps];
    float s = 0.f;
    if ((int)blockIdx.x < dense_blocks) {
        const long long n4 = n >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(grad);
        for (long long i = (long long)blockIdx.x * kUpdThreads + threadIdx.x; i < n4;
             i += (long long)dense_blocks * kUpdThreads) {
            const float4 v = g4[i];
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = grad[(n4 << 2) + threadIdx.x]; s += v * v; }
    } else {
        // one warp per token slot (grid-stride): thousands of independent {word -> owner -> row} chains in flight
        const int row_blocks = gridDim.x - dense_blocks;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int row = ((int)blockIdx.x - dense_blocks) * kUpdWarps + warp; row < n_rows;
             row += row_blocks * kUpdWarps) {
            const long long w = words[row];
            if (w == 0 || w >= topn || owner[w] != row) continue;     // warp-uniform
            const float* gr = g_emb + (size_t)w * E;
            if ((E & 3) == 0) {
                const float4* g4 = reinterpret_cast<const float4*>(gr);
                float4 v[kRowVec];
#pragma unroll
                for (int i = 0; i < kRowVec; ++i)
                    v[i] = lane + 32 * i < (E >> 2) ? g4[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < kRowVec; ++i) s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
                for (int c = lane + 32 * kRowVec; c < (E >> 2); c += 32) {
                    const float4 u = g4[c];
                    s += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
                }
            } else {
                for (int c = lane; c < E; c += 32) { const float u = gr[c]; s += u * u; }
            }
        }
    }
    const float t = block_sum(s, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kUpdThreads)
update_apply_kernel(float* __restrict__ param, float* __restrict__ grad, long long n,
                    const long long* __restrict__ words, int* __restrict__ owner, float* __restrict__ g_emb,
                    float* __restrict__ emb_w, int n_rows, int E, int topn, int dense_blocks,
                    const float* __restrict__ partials, int n_partials, float max_norm, float lr, float grad_scale,
                    float* __restrict__ total_norm, unsigned long long* __restrict__ step_counter) {
    GPT_PDL_ENTER();
    __shared__ float s_red[kUpdWarps];
    __shared__ float s_coef;
    if (blockIdx.x == 0 && threadIdx.x == 0 && step_counter != nullptr) *step_counter += 1ull;
    {   // same order in every CTA => every CTA sees bit-identical coef
        float s = 0.f;
        for (int i = threadIdx.x; i < n_partials; i += kUpdThreads) s += partials[i];
        const float t = block_sum(s, s_red);
        if (threadIdx.x == 0) {
            const float norm = sqrtf(t) * grad_scale;
            s_coef = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
            if (blockIdx.x == 0 && total_norm != nullptr) *total_norm = norm;
        }
        __syncthreads();
    }
    const float a = lr * s_coef * grad_scale;
    if ((int)blockIdx.x < dense_blocks) {
        const long long n4 = n >> 2;
        float4* g4 = reinterpret_cast<float4*>(grad);
        float4* p4 = reinterpret_cast<float4*>(param);
        for (long long i = (long long)blockIdx.x * kUpdThreads + threadIdx.x; i < n4;
             i += (long long)dense_blocks * kUpdThreads) {
            const float4 g = g4[i];
            float4 p = p4[i];
            p.x -= a * g.x; p.y -= a * g.y; p.z -= a * g.z; p.w -= a * g.w;
            p4[i] = p;
            g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
            const long long i = (n4 << 2) + threadIdx.x;
            param[i] -= a * grad[i];
            grad[i] = 0.f;
        }
    } else {
        const int row_blocks = gridDim.x - dense_blocks;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int row = ((int)blockIdx.x - dense_blocks) * kUpdWarps + warp; row < n_rows;
             row += row_blocks * kUpdWarps) {
            const long long w = words[row];
            if (w == 0 || w >= topn || owner[w] != row) continue;     // warp-uniform
            float* gr = g_emb + (size_t)w * E;
            float* wr = emb_w + (size_t)w * E;
            if ((E & 3) == 0) {
                float4* g4 = reinterpret_cast<float4*>(gr);
                float4* w4 = reinterpret_cast<float4*>(wr);
                float4 gv[kRowVec], wv[kRowVec];
#pragma unroll
                for (int i = 0; i < kRowVec; ++i) {      // all loads of the row first: one round trip to HBM
                    const bool in = lane + 32 * i < (E >> 2);
                    gv[i] = in ? g4[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    wv[i] = in ? w4[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < kRowVec; ++i) {
                    if (lane + 32 * i < (E >> 2)) {
                        wv[i].x -= a * gv[i].x; wv[i].y -= a * gv[i].y; wv[i].z -= a * gv[i].z; wv[i].w -= a * gv[i].w;
                        w4[lane + 32 * i] = wv[i];
                        g4[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                for (int c = lane + 32 * kRowVec; c < (E >> 2); c += 32) {
                    const float4 g = g4[c];
                    float4 q = w4[c];
                    q.x -= a * g.x; q.y -= a * g.y; q.z -= a * g.z; q.w -= a * g.w;
                    w4[c] = q;
                    g4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                for (int c = lane; c < E; c += 32) {
                    wr[c] -= a * gr[c];
                    gr[c] = 0.f;
                }
            }
            __syncwarp();
            if (lane == 0) owner[w] = 0x7fffffff;   // readers of other tokens of this word see "not mine" either way
        }
    }
}

int plan(long long n, int n_rows, int* dense_blocks, int* row_blocks) {
    long long db = (n / 4 + kUpdThreads * 4 - 1) / (kUpdThreads * 4);     // ~4 float4 per thread
    if (db < 1) db = 1;
    if (db > 296) db = 296;
    int rb = n_rows > 0 ? (n_rows + kUpdWarps - 1) / kUpdWarps : 0;     // one warp per token slot ...
    if (rb > kMaxPartials - 296) rb = kMaxPartials - 296;               // ... grid-stride beyond that
    *dense_blocks = n > 0 ? (int)db : 0;
    *row_blocks = rb;
    return *dense_blocks + rb;
}

}  // namespace

extern "C" int gpt_update_partials(long long n, int n_rows) {
    int d, r;
    return plan(n, n_rows, &d, &r);
}

extern "C" int gpt_update_sqnorm(const float* grad, long long n, const int64_t* words, const int32_t* owner,
                                 const float* g_emb, int n_rows, int E, int topn, float* partials, void* stream) {
    GPT_CHECK_ARG(n >= 0 && n_rows >= 0 && partials);
    GPT_CHECK_ARG(n == 0 || (grad && (reinterpret_cast<uintptr_t>(grad) & 15) == 0));
    GPT_CHECK_ARG(n_rows == 0 || (words && owner && g_emb && E >= 1));
    int d, r;
    const int blocks = plan(n, n_rows, &d, &r);
    if (blocks == 0) return GPT_OK;
    if (blocks > kMaxPartials) return GPT_ERR_UNSUPPORTED;
    gpt_launch(update_sqnorm_kernel, dim3(blocks), dim3(kUpdThreads), 0, (cudaStream_t)stream, 
        grad, n, reinterpret_cast<const long long*>(words), owner, g_emb, n_rows, E, topn, d, partials);
    return gpt_launch_status();
}

extern "C" int gpt_update_apply(float* param, float* grad, long long n, const int64_t* words, int32_t* owner,
                                float* g_emb, float* emb_w, int n_rows, int E, int topn, const float* partials,
                                float max_norm, float lr, float grad_scale, float* total_norm,
                                uint64_t* step_counter, void* stream) {
    GPT_CHECK_ARG(n >= 0 && n_rows >= 0 && partials);
    GPT_CHECK_ARG(n == 0 || (param && grad && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
                             (reinterpret_cast<uintptr_t>(param) & 15) == 0));
    GPT_CHECK_ARG(n_rows == 0 || (words && owner && g_emb && emb_w && E >= 1));
    int d, r;
    const int blocks = plan(n, n_rows, &d, &r);
    if (blocks == 0) return GPT_OK;
    gpt_launch(update_apply_kernel, dim3(blocks), dim3(kUpdThreads), 0, (cudaStream_t)stream, 
        param, grad, n, reinterpret_cast<const long long*>(words), owner, g_emb, emb_w, n_rows, E, topn, d, partials,
        blocks, max_norm, lr, grad_scale, total_norm, reinterpret_cast<unsigned long long*>(step_counter));
    return gpt_launch_status();
}
