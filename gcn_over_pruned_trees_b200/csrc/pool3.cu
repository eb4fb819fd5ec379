// K4 -- fused sentence / subject / object masked pooling.
//
// Replaces the three pool() passes + cat of /root/reference/model/gcn.py:116-121 (pool: gcn.py:473-483):
//   h_out    = pool(h, not-in-tree mask)   mask = (rowsum + colsum) == 0            gcn.py:262
//   subj_out = pool(h, subj_pos != 0)      obj_out = pool(h, obj_pos != 0)          gcn.py:116
//   out      = cat[h_out, subj_out, obj_out]  -> [B, 3H]
// 'max' fills masked positions with -1e12 (utils/constant.py:35), so a fully masked pool yields -1e12;
// 'avg' divides by (T - #masked), 'sum' just adds.  h is read once for all three pools.
//
// One thread per (sentence, column); threads of a warp read consecutive columns (coalesced), flags are staged
// in shared memory once per CTA.
#include "gpt_common.cuh"

namespace {

constexpr int kPoolThreads = 128;
constexpr float kNegFill = -1e12f;
enum { POOL_MAX = 0, POOL_AVG = 1, POOL_SUM = 2 };

__global__ void __launch_bounds__(kPoolThreads)
pool3_fwd_kernel(const float* __restrict__ h, const unsigned char* __restrict__ flags, int T, int H, int type,
                 float* __restrict__ out, int* __restrict__ argmax) {
    extern __shared__ unsigned char s_flags[];
    const int b = blockIdx.y;
    for (int t = threadIdx.x; t < T; t += kPoolThreads) s_flags[t] = flags[(size_t)b * T + t];
    __syncthreads();
    const int c = blockIdx.x * kPoolThreads + threadIdx.x;
    if (c >= H) return;
    const float* hb = h + (size_t)b * T * H + c;
    float acc[3];
    int arg[3] = {-1, -1, -1}, cnt[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = (type == POOL_MAX) ? kNegFill : 0.f;
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
        const unsigned f = s_flags[t];
        if (f == 0) continue;  // CTA-uniform: token is in none of the three pools
        const float v = hb[(size_t)t * H];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (f & (1u << k)) {
                ++cnt[k];
                if (type == POOL_MAX) {
                    if (v > acc[k]) { acc[k] = v; arg[k] = t; }
                } else {
                    acc[k] += v;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float r = acc[k];
        if (type == POOL_AVG) r = r / (float)cnt[k];  // 0/0 -> nan, as h.sum(1) / (T - mask.sum(1)) does
        out[(size_t)b * 3 * H + (size_t)k * H + c] = r;
        if (argmax) argmax[(size_t)b * 3 * H + (size_t)k * H + c] = arg[k];
    }
}

__global__ void __launch_bounds__(kPoolThreads)
pool3_bwd_kernel(const float* __restrict__ gout, const int* __restrict__ argmax,
                 const unsigned char* __restrict__ flags, int T, int H, int type, float* __restrict__ dh) {
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
    if (B > 65535 || T > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    dim3 grid((H + kPoolThreads - 1) / kPoolThreads, B);
    pool3_fwd_kernel<<<grid, kPoolThreads, T, (cudaStream_t)stream>>>(h, flags, T, H, pool_type, out, argmax);
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
    pool3_bwd_kernel<<<grid, kPoolThreads, T, (cudaStream_t)stream>>>(gout, argmax, flags, T, H, pool_type, dh);
    return gpt_launch_status();
}
