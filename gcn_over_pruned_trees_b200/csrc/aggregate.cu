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
// Mapping: one CTA per (sentence, 32*VEC-column slice).  The slice of all T projected rows of the sentence is
// staged in shared memory with cp.async (each row is read from HBM exactly once), the sentence's CSR is staged
// next to it, then one warp per node gathers its <= deg+1 rows out of shared memory with 32*VEC-wide loads and
// applies the epilogue.  HBM traffic = read y once + write out once + CSR: the algorithmic minimum.
//
// Backward (adjacency is symmetric, so A^T = A and the same CSR is reused):
//   g_i  = gout_i * dropscale * [out_i > 0] / denom_i          (staged into shared memory)
//   dy_j = g_j + sum_{i in row j} g_i ,   dbias = 2 * sum_i g_i
#include "gpt_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowBlock = 8;  // rows that share one Philox call per column

struct AggParams {
    const float* y;      // [B*T, H] projected rows (fwd) / gout (bwd)
    const float* aux;    // bwd: out of the forward pass
    const int* rowptr;   // [B, T+1]
    const int* col;      // [B, cap]
    const float* denom;  // [B*T]
    const unsigned char* flags;  // [B*T]
    const float* bias;   // [H] (fwd)
    float* out;          // [B*T, H]
    float* dbias;        // [H] (bwd, atomically accumulated)
    const float* drop_mask;              // optional explicit, pre-scaled mask [B*T, H]
    const unsigned long long* rng;       // optional {seed, step} on the device
    int B, T, H, cap, use_adj;
    unsigned subseq, thresh16;           // dropout: keep iff rand16 >= thresh16
    float drop_scale;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int VEC>
struct Vec;
template <>
struct Vec<1> { using type = float; };
template <>
struct Vec<2> { using type = float2; };
template <>
struct Vec<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void lds_add(float (&acc)[VEC], const float* p) {
    typename Vec<VEC>::type v = *reinterpret_cast<const typename Vec<VEC>::type*>(p);
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += f[i];
}

template <int VEC, bool ALIGNED>
__device__ __forceinline__ void store_row(float* dst, const float (&v)[VEC], int c, int H) {
    if (ALIGNED) {
        if (c < H) {  // H % 4 == 0 and c % VEC == 0: the vector is entirely inside or outside
            typename Vec<VEC>::type pack;
            float* f = reinterpret_cast<float*>(&pack);
#pragma unroll
            for (int i = 0; i < VEC; ++i) f[i] = v[i];
            *reinterpret_cast<typename Vec<VEC>::type*>(dst) = pack;
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            if (c + i < H) dst[i] = v[i];
    }
}

// Stage the sentence's CSR ([T+1] offsets + nnz columns) into shared memory.
__device__ __forceinline__ void stage_csr(const AggParams& p, int b, int* s_rp, int* s_col) {
    const int* rp = p.rowptr + (size_t)b * (p.T + 1);
    for (int t = threadIdx.x; t <= p.T; t += kThreads) s_rp[t] = rp[t];
    const int nnz = p.use_adj ? min(rp[p.T], p.cap) : 0;
    const int* cb = p.col + (size_t)b * p.cap;
    for (int e = threadIdx.x; e < nnz; e += kThreads) s_col[e] = cb[e];
}

template <int VEC, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) aggregate_fwd_kernel(const AggParams p) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int HS = 32 * VEC;
    const int T = p.T, H = p.H;
    const int b = blockIdx.y, col0 = blockIdx.x * HS;
    float* tile = smem_f;
    int* s_rp = reinterpret_cast<int*>(tile + (size_t)T * HS);
    int* s_col = s_rp + (T + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage y[b, :, col0 : col0+HS] -----------------------------------------------------------------------
    const float* yb = p.y + (size_t)b * T * H;
    if (ALIGNED) {
        constexpr int CPR = HS / 4;  // 16-byte chunks per row
        for (int q = tid; q < T * CPR; q += kThreads) {
            const int row = q / CPR, cc = (q % CPR) * 4, c = col0 + cc;
            float* dst = tile + row * HS + cc;
            if (c < H) cp_async16(dst, yb + (size_t)row * H + c);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int q = tid; q < T * HS; q += kThreads) {
            const int row = q / HS, cc = q % HS, c = col0 + cc;
            if (c < H) cp_async4(tile + q, yb + (size_t)row * H + c);
            else tile[q] = 0.f;
        }
    }
    stage_csr(p, b, s_rp, s_col);
    cp_async_wait_all();
    __syncthreads();

    // ---- per-thread constants ---------------------------------------------------------------------------------
    const int c_lane = col0 + lane * VEC;
    float bias2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) bias2[v] = (c_lane + v < H) ? 2.0f * p.bias[c_lane + v] : 0.f;
    const bool philox = (p.rng != nullptr) && (p.thresh16 > 0);
    unsigned long long seed = 0, step = 0;
    if (philox) { seed = p.rng[0]; step = p.rng[1]; }

    // ---- one warp per node, nodes taken in blocks of kRowBlock consecutive rows -------------------------------
    for (int blk = warp; blk * kRowBlock < T; blk += kWarps) {
        unsigned long long rlo[VEC], rhi[VEC];  // 8 x 16 random bits per column: one per row of the block
#pragma unroll
        for (int v = 0; v < VEC; ++v) { rlo[v] = 0; rhi[v] = 0; }
        if (philox) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const Philox4 q = philox4x32_10((uint32_t)(c_lane + v) | (p.subseq << 20), (uint32_t)blk, (uint32_t)b,
                                                (uint32_t)step, (uint32_t)seed,
                                                (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
                rlo[v] = (unsigned long long)q.x | ((unsigned long long)q.y << 32);
                rhi[v] = (unsigned long long)q.z | ((unsigned long long)q.w << 32);
            }
        }
#pragma unroll 2
        for (int r = 0; r < kRowBlock; ++r) {
            const int i = blk * kRowBlock + r;
            if (i >= T) continue;
            const size_t grow = (size_t)b * T + i;
            float res[VEC];
            if (p.flags[grow] != 0) {  // warp-uniform: row is observable (in tree, or an entity token)
                float acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
                lds_add<VEC>(acc, tile + i * HS + lane * VEC);  // the separate W(h) self term
                if (p.use_adj) {
                    const int e1 = s_rp[i + 1];
                    for (int e = s_rp[i]; e < e1; ++e) lds_add<VEC>(acc, tile + s_col[e] * HS + lane * VEC);
                }
                const float dn = p.denom[grow];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float z = (acc[v] + bias2[v]) / dn;
                    z = fmaxf(z, 0.f);
                    if (philox) {
                        const uint32_t bits = (uint32_t)(((r < 4 ? rlo[v] : rhi[v]) >> ((r & 3) * 16)) & 0xffffull);
                        z = (bits >= p.thresh16) ? z * p.drop_scale : 0.f;
                    }
                    res[v] = z;
                }
                if (p.drop_mask != nullptr) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        if (c_lane + v < H) res[v] *= p.drop_mask[grow * H + c_lane + v];
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) res[v] = 0.f;
            }
            store_row<VEC, ALIGNED>(p.out + grow * H + c_lane, res, c_lane, H);
        }
    }
}

template <int VEC, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) aggregate_bwd_kernel(const AggParams p) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int HS = 32 * VEC;
    const int T = p.T, H = p.H;
    const int b = blockIdx.y, col0 = blockIdx.x * HS;
    float* tile = smem_f;
    int* s_rp = reinterpret_cast<int*>(tile + (size_t)T * HS);
    int* s_col = s_rp + (T + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage g = gout * dropscale * [out > 0] / denom --------------------------------------------------------
    const size_t base = (size_t)b * T * H;
    const float* gb = p.y + base;
    const float* ob = p.aux + base;
    const float* mb = p.drop_mask ? p.drop_mask + base : nullptr;
    const float* dnb = p.denom + (size_t)b * T;
    if (ALIGNED) {
        constexpr int CPR = HS / 4;
        for (int q = tid; q < T * CPR; q += kThreads) {
            const int row = q / CPR, cc = (q % CPR) * 4, c = col0 + cc;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < H) {
                const size_t off = (size_t)row * H + c;
                const float4 go = *reinterpret_cast<const float4*>(gb + off);
                const float4 o = *reinterpret_cast<const float4*>(ob + off);
                float4 m = make_float4(p.drop_scale, p.drop_scale, p.drop_scale, p.drop_scale);
                if (mb) m = *reinterpret_cast<const float4*>(mb + off);
                const float dn = dnb[row];
                g.x = o.x > 0.f ? go.x * m.x / dn : 0.f;
                g.y = o.y > 0.f ? go.y * m.y / dn : 0.f;
                g.z = o.z > 0.f ? go.z * m.z / dn : 0.f;
                g.w = o.w > 0.f ? go.w * m.w / dn : 0.f;
            }
            *reinterpret_cast<float4*>(tile + row * HS + cc) = g;
        }
    } else {
        for (int q = tid; q < T * HS; q += kThreads) {
            const int row = q / HS, cc = q % HS, c = col0 + cc;
            float g = 0.f;
            if (c < H) {
                const size_t off = (size_t)row * H + c;
                const float m = mb ? mb[off] : p.drop_scale;
                g = ob[off] > 0.f ? gb[off] * m / dnb[row] : 0.f;
            }
            tile[q] = g;
        }
    }
    stage_csr(p, b, s_rp, s_col);
    __syncthreads();

    // ---- dy_j = g_j + sum_{i in row j} g_i ; column sums for dbias ---------------------------------------------
    const int c_lane = col0 + lane * VEC;
    float csum[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) csum[v] = 0.f;
    for (int j = warp; j < T; j += kWarps) {
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        lds_add<VEC>(acc, tile + j * HS + lane * VEC);
#pragma unroll
        for (int v = 0; v < VEC; ++v) csum[v] += acc[v];
        if (p.use_adj) {
            const int e1 = s_rp[j + 1];
            for (int e = s_rp[j]; e < e1; ++e) lds_add<VEC>(acc, tile + s_col[e] * HS + lane * VEC);
        }
        store_row<VEC, ALIGNED>(p.out + ((size_t)b * T + j) * H + c_lane, acc, c_lane, H);
    }
    if (p.dbias != nullptr) {
        __syncthreads();  // everyone is done reading the tile; reuse its head for the cross-warp reduction
        float* red = tile;  // [kWarps][HS]
#pragma unroll
        for (int v = 0; v < VEC; ++v) red[warp * HS + lane * VEC + v] = csum[v];
        __syncthreads();
        for (int c = tid; c < HS; c += kThreads) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += red[w * HS + c];
            if (col0 + c < H) atomicAdd(p.dbias + col0 + c, 2.0f * s);  // the bias enters the layer twice
        }
    }
}

size_t agg_smem_bytes(int T, int vec) {
    // tile [max(T, kWarps)][32*vec] floats + rowptr [T+1] + col [3T]
    const size_t rows = (size_t)(T > kWarps ? T : kWarps);
    return rows * 32 * vec * sizeof(float) + (size_t)(T + 1 + 3 * T) * sizeof(int);
}

// widest slice that still leaves >= 3 CTAs per SM and fills the machine at least ~2 waves
int pick_vec(int B, int T, int H, int force) {
    if (force == 1 || force == 2 || force == 4) return force;
    int best = 1;
    for (int vec = 4; vec >= 1; vec >>= 1) {
        if (agg_smem_bytes(T, vec) > 75 * 1024) continue;  // 3 CTAs per SM
        const long slices = (H + 32 * vec - 1) / (32 * vec);
        if (vec > 1 && slices * B < 2 * 148) continue;
        best = vec;
        break;
    }
    return best;
}

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t a = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (a != cudaSuccess) return (int)a;
    }
    return GPT_OK;
}

template <int VEC, bool ALIGNED>
int launch(bool fwd, const AggParams& p, cudaStream_t st) {
    const size_t smem = agg_smem_bytes(p.T, VEC);
    if (smem > 224 * 1024) return GPT_ERR_UNSUPPORTED;
    dim3 grid((p.H + 32 * VEC - 1) / (32 * VEC), p.B);
    int rc;
    if (fwd) {
        if ((rc = ensure_smem(aggregate_fwd_kernel<VEC, ALIGNED>, smem)) != GPT_OK) return rc;
        aggregate_fwd_kernel<VEC, ALIGNED><<<grid, kThreads, smem, st>>>(p);
    } else {
        if ((rc = ensure_smem(aggregate_bwd_kernel<VEC, ALIGNED>, smem)) != GPT_OK) return rc;
        aggregate_bwd_kernel<VEC, ALIGNED><<<grid, kThreads, smem, st>>>(p);
    }
    return gpt_launch_status();
}

int dispatch(bool fwd, const AggParams& p, int force_vec, cudaStream_t st) {
    const bool aligned = (p.H % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                         (p.aux == nullptr || (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) &&
                         (p.drop_mask == nullptr || (reinterpret_cast<uintptr_t>(p.drop_mask) & 15) == 0);
    int vec = pick_vec(p.B, p.T, p.H, force_vec);
    while (vec > 1 && agg_smem_bytes(p.T, vec) > 224 * 1024) vec >>= 1;
    if (aligned) {
        if (vec == 4) return launch<4, true>(fwd, p, st);
        if (vec == 2) return launch<2, true>(fwd, p, st);
        return launch<1, true>(fwd, p, st);
    }
    if (vec == 4) return launch<4, false>(fwd, p, st);
    if (vec == 2) return launch<2, false>(fwd, p, st);
    return launch<1, false>(fwd, p, st);
}

}  // namespace

extern "C" int gpt_gcn_aggregate_fwd(const float* y, const int32_t* rowptr, const int32_t* col, const float* denom,
                                     const uint8_t* flags, const float* bias, float* out, int B, int T, int H,
                                     int use_adj, float drop_p, const uint64_t* rng_state, uint32_t subseq,
                                     const float* drop_mask, int force_vec, void* stream) {
    GPT_CHECK_ARG(y && rowptr && col && denom && flags && bias && out);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && H < (1 << 20) && drop_p >= 0.f && drop_p < 1.f);
    GPT_CHECK_ARG(!(drop_p > 0.f && rng_state == nullptr));  // dropout needs the device-side {seed, step}
    if (B == 0) return GPT_OK;
    if (B > 65535) return GPT_ERR_UNSUPPORTED;
    AggParams p{};
    p.y = y; p.rowptr = rowptr; p.col = col; p.denom = denom; p.flags = flags; p.bias = bias; p.out = out;
    p.drop_mask = drop_mask;
    p.rng = (drop_p > 0.f) ? reinterpret_cast<const unsigned long long*>(rng_state) : nullptr;
    p.B = B; p.T = T; p.H = H; p.cap = 3 * T; p.use_adj = use_adj;
    p.subseq = subseq & 0xfffu;
    unsigned th = (unsigned)(drop_p * 65536.0f + 0.5f);
    p.thresh16 = th > 65535u ? 65535u : th;
    p.drop_scale = (p.thresh16 > 0) ? 65536.0f / (65536.0f - (float)p.thresh16) : 1.0f;
    return dispatch(true, p, force_vec, (cudaStream_t)stream);
}

extern "C" int gpt_gcn_aggregate_bwd(const float* gout, const float* out, const int32_t* rowptr, const int32_t* col,
                                     const float* denom, float* dy, float* dbias, int B, int T, int H, int use_adj,
                                     float drop_p, const float* drop_mask, int force_vec, void* stream) {
    GPT_CHECK_ARG(gout && out && rowptr && col && denom && dy);
    GPT_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && drop_p >= 0.f && drop_p < 1.f);
    if (B == 0) return GPT_OK;
    if (B > 65535) return GPT_ERR_UNSUPPORTED;
    AggParams p{};
    p.y = gout; p.aux = out; p.rowptr = rowptr; p.col = col; p.denom = denom; p.out = dy; p.dbias = dbias;
    p.drop_mask = drop_mask;
    p.B = B; p.T = T; p.H = H; p.cap = 3 * T; p.use_adj = use_adj;
    unsigned th = (unsigned)(drop_p * 65536.0f + 0.5f);
    th = th > 65535u ? 65535u : th;
    // same scale the forward applied to kept elements (dropped ones have out == 0 and are masked by [out > 0])
    p.drop_scale = (th > 0 && drop_mask == nullptr) ? 65536.0f / (65536.0f - (float)th) : 1.0f;
    return dispatch(false, p, force_vec, (cudaStream_t)stream);
}
