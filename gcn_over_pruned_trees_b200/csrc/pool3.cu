// K4 -- fused sentence / subject / object masked pooling.
//
// Replaces the three pool() passes + cat of /root/reference/model/gcn.py:116-121 (pool: gcn.py:473-483):
//   h_out    = pool(h, not-in-tree mask)   mask = (rowsum + colsum) == 0            gcn.py:262
//   subj_out = pool(h, subj_pos != 0)      obj_out = pool(h, obj_pos != 0)          gcn.py:116
//   out      = cat[h_out, subj_out, obj_out]  -> [B, 3H]
// 'max' fills masked positions with -1e12 (utils/constant.py:35), so a fully masked pool yields -1e12;
// 'avg' divides by (T - #masked), 'sum' just adds.  h is read once for all three pools.
//
// Forward: one CTA per sentence.  Warp 0 first compacts the tokens that are in at least one pool (flags != 0; at
// prune_k = 1 that is ~20 of up to 96) into a shared list, in token order.  The CTA then splits into G groups of
// H/4 threads: a thread owns one 128-bit column and every G-th listed token, so all of its loads are independent
// and in flight together; the G partial (value, argmax) triples meet in shared memory.  Ties keep the smallest
// token index, whatever G is.
#include "gpt_common.cuh"

namespace {

constexpr int kPoolThreads = 128;      // backward
constexpr int kPoolFwdThreads = 256;
constexpr float kNegFill = -1e12f;
enum { POOL_MAX = 0, POOL_AVG = 1, POOL_SUM = 2 };

template <int V>   // V = 4: H % 4 == 0, float4 columns; V = 1: any H
__global__ void __launch_bounds__(kPoolFwdThreads)
pool3_fwd_kernel(const float* __restrict__ h, const unsigned char* __restrict__ flags, int T, int H, int type,
                 float* __restrict__ out, int* __restrict__ argmax) {
    GPT_PDL_ENTER();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_n, s_cnt[3];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int cols = H / V;                                   // column vectors per row
    const int G = cols < kPoolFwdThreads ? kPoolFwdThreads / cols : 1;
    int* s_list = reinterpret_cast<int*>(smem_raw);           // [T]  t | flags << 16
    float* s_val = reinterpret_cast<float*>(s_list + ((T + 3) & ~3));   // [G][3][H]
    int* s_arg = reinterpret_cast<int*>(s_val + (size_t)G * 3 * H);     // [G][3][H]
    // the sentence's flags, fetched by the whole CTA at once (one exposed global latency instead of one per 32 tokens
    // in warp 0's compaction loop below); parked in the first partial-result slot, which is not written before the
    // second barrier
    unsigned char* s_flags = reinterpret_cast<unsigned char*>(s_val);
    for (int t = tid; t < T; t += kPoolFwdThreads) s_flags[t] = flags[(size_t)b * T + t];
    __syncthreads();
    if (tid < 32) {
        int n = 0, c0 = 0, c1 = 0, c2 = 0;
        for (int base = 0; base < T; base += 32) {
            const int t = base + tid;
            const unsigned f = t < T ? s_flags[t] : 0u;
            const unsigned m = __ballot_sync(GPT_FULL_MASK, f != 0);
            if (f != 0) s_list[n + __popc(m & ((1u << tid) - 1u))] = t | (int)(f << 16);
            n += __popc(m);
            c0 += __popc(__ballot_sync(GPT_FULL_MASK, f & 1u));
            c1 += __popc(__ballot_sync(GPT_FULL_MASK, f & 2u));
            c2 += __popc(__ballot_sync(GPT_FULL_MASK, f & 4u));
        }
        if (tid == 0) { s_n = n; s_cnt[0] = c0; s_cnt[1] = c1; s_cnt[2] = c2; }
    }
    __syncthreads();
    const int n = s_n;
    const float init = type == POOL_MAX ? kNegFill : 0.f;
    const float* hb = h + (size_t)b * T * H;
    for (int item = tid; item < G * cols; item += kPoolFwdThreads) {
        const int g = item / cols, c = item - g * cols;
        float acc[3][V];
        int arg[3][V];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int v = 0; v < V; ++v) { acc[k][v] = init; arg[k][v] = -1; }
#pragma unroll 4
        for (int i = g; i < n; i += G) {
            const int e = s_list[i];
            const int t = e & 0xffff;
            const unsigned f = (unsigned)e >> 16;
            float x[V];
            if (V == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(hb + (size_t)t * H) + c);
                x[0] = q.x; x[1 % V] = q.y; x[2 % V] = q.z; x[3 % V] = q.w;
            } else {
                x[0] = __ldg(hb + (size_t)t * H + c);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (f & (1u << k)) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (type == POOL_MAX) {
                            if (x[v] > acc[k][v]) { acc[k][v] = x[v]; arg[k][v] = t; }
                        } else {
                            acc[k][v] += x[v];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int v = 0; v < V; ++v) {
                s_val[((size_t)g * 3 + k) * H + c * V + v] = acc[k][v];
                s_arg[((size_t)g * 3 + k) * H + c * V + v] = arg[k][v];
            }
    }
    __syncthreads();
    for (int i = tid; i < 3 * H; i += kPoolFwdThreads) {      // i = k * H + column
        float r = s_val[i];
        int a = s_arg[i];
        for (int g = 1; g < G; ++g) {
            const float v = s_val[(size_t)g * 3 * H + i];
            const int av = s_arg[(size_t)g * 3 * H + i];
            if (type == POOL_MAX) {
                if (v > r || (v == r && av >= 0 && (a < 0 || av < a))) { r = v; a = av; }
            } else {
                r += v;
            }
        }
        if (type == POOL_AVG) r = r / (float)s_cnt[i / H];    // 0/0 -> nan, as h.sum(1) / (T - mask.sum(1)) does
        out[(size_t)b * 3 * H + i] = r;
        if (argmax) argmax[(size_t)b * 3 * H + i] = a;
    }
}

size_t pool_fwd_smem(int T, int H, int V) {
    const int cols = H / V;
    const int G = cols < kPoolFwdThreads ? kPoolFwdThreads / cols : 1;
    const size_t partials = (size_t)G * 3 * H * 8;      // the flags (T bytes) are staged in the same space first
    return (size_t)((T + 3) & ~3) * 4 + (partials > (size_t)T ? partials : (size_t)((T + 15) & ~15));
}

__global__ void __launch_bounds__(kPoolThreads)
pool3_bwd_kernel(const float* __restrict__ gout, const int* __restrict__ argmax,
                 const unsigned char* __restrict__ flags, int T, int H, int type, float* __restrict__ dh,
                 const uint32_t* __restrict__ act, const float* __restrict__ denom, float scale) {
    GPT_PDL_ENTER();
    extern __shared__ unsigned char s_flags[];
    const int b = blockIdx.y;
    int cnt[3] = {0, 0, 0};
    for (int t = threadIdx.x; t < T; t += kPoolThreads) s_flags[t] = flags[(size_t)b * T + t];
    __syncthreads();
    const int c = blockIdx.x * kPoolThreads + threadIdx.x;
    if (c >= H) return;
    float g[3];
    int arg[3] = {-1, -1, -1};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        g[k] = gout[(size_t)b * 3 * H + (size_t)k * H + c];
        if (type == POOL_MAX) arg[k] = argmax[(size_t)b * 3 * H + (size_t)k * H + c];
    }
    if (type == POOL_AVG) {
        for (int t = 0; t < T; ++t) {
            const unsigned f = s_flags[t];
#pragma unroll
            for (int k = 0; k < 3; ++k) cnt[k] += (f >> k) & 1u;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = g[k] / (float)cnt[k];
    }
    float* db = dh + (size_t)b * T * H + c;
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
        float v = 0.f;
        if (type == POOL_MAX) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (arg[k] == t) v += g[k];
        } else {
            const unsigned f = s_flags[t];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (f & (1u << k)) v += g[k];
        }
        if (act != nullptr) {   // fused first step of K2's backward: g = dh * dropscale * [out > 0] / denom
            const uint32_t w = act[((size_t)b * ((H + 31) / 32) + (c >> 5)) * T + t];
            v = v * ((float)((w >> (c & 31)) & 1u) * scale) * __frcp_rn(denom[(size_t)b * T + t]);
        }
        db[(size_t)t * H] = v;
    }
}

}  // namespace

extern "C" int gpt_pool3_fwd(const float* h, const uint8_t* flags, int B, int T, int H, int pool_type, float* out,
                             int32_t* argmax, void* stream) {
    GPT_CHECK_ARG(h && flags && out);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && pool_type >= 0 && pool_type <= 2);
    GPT_CHECK_ARG(pool_type != POOL_MAX || argmax != nullptr);
    if (B == 0) return GPT_OK;
    if (T > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    const bool vec = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(h) & 15) == 0);
    const size_t smem = pool_fwd_smem(T, H, vec ? 4 : 1);
    if (smem > 200 * 1024) return GPT_ERR_UNSUPPORTED;
    if (int e = vec ? gpt_smem_opt_in(pool3_fwd_kernel<4>, smem) : gpt_smem_opt_in(pool3_fwd_kernel<1>, smem)) return e;
    if (vec) gpt_launch(pool3_fwd_kernel<4>, dim3(B), dim3(kPoolFwdThreads), smem, (cudaStream_t)stream, h, flags, T, H, pool_type, out, argmax);
    else gpt_launch(pool3_fwd_kernel<1>, dim3(B), dim3(kPoolFwdThreads), smem, (cudaStream_t)stream, h, flags, T, H, pool_type, out, argmax);
    return gpt_launch_status();
}

extern "C" int gpt_pool3_bwd(const float* gout, const int32_t* argmax, const uint8_t* flags, int B, int T, int H,
                             int pool_type, float* dh, void* stream) {
    GPT_CHECK_ARG(gout && flags && dh);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && pool_type >= 0 && pool_type <= 2);
    GPT_CHECK_ARG(pool_type != POOL_MAX || argmax != nullptr);
    if (B == 0) return GPT_OK;
    if (B > 65535 || T > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    dim3 grid((H + kPoolThreads - 1) / kPoolThreads, B);
    gpt_launch(pool3_bwd_kernel, grid, dim3(kPoolThreads), (size_t)T, (cudaStream_t)stream, gout, argmax, flags, T, H, pool_type, dh, nullptr,
                                                                      nullptr, 1.f);
    return gpt_launch_status();
}

extern "C" int gpt_pool3_bwd_masked(const float* gout, const int32_t* argmax, const uint8_t* flags,
                                    const uint32_t* act, const float* denom, float drop_scale, int B, int T, int H,
                                    int pool_type, float* g, void* stream) {
    GPT_CHECK_ARG(gout && flags && g && act && denom);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && pool_type >= 0 && pool_type <= 2);
    GPT_CHECK_ARG(pool_type != POOL_MAX || argmax != nullptr);
    if (B == 0) return GPT_OK;
    if (B > 65535 || T > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    dim3 grid((H + kPoolThreads - 1) / kPoolThreads, B);
    gpt_launch(pool3_bwd_kernel, grid, dim3(kPoolThreads), (size_t)T, (cudaStream_t)stream, gout, argmax, flags, T, H, pool_type, g, act,
                                                                      denom, drop_scale);
    return gpt_launch_status();
}
