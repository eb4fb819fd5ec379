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
// Mapping: one CTA per (sentence, HS-column slice), HS = 4*LPR in {32, 64, 128}.  The slice of all T projected
// rows of the sentence is staged in shared memory with cp.async (each row is read from HBM exactly once), the
// sentence's CSR / denom / flags are staged next to it (16-bit indices), then LPR lanes per node gather the
// node's <= deg+1 rows out of shared memory with 128-bit loads (a warp works on 32/LPR nodes at a time, two
// rows in flight per lane group) and apply the epilogue.  HBM traffic = read y once + write out once + CSR.
//
// Backward (adjacency is symmetric, so A^T = A and the same CSR is reused):
//   g_i  = gout_i * dropscale * [out_i > 0] / denom_i          (staged into shared memory)
//   dy_j = g_j + sum_{i in row j} g_i ,   dbias = 2 * sum_i g_i
#include "gpt_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowBlock = 8;  // consecutive rows that share one Philox call per column

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

__device__ __forceinline__ void add4(float4& a, const float4 b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

// Shared-memory accessors on 32-bit shared-window addresses: keeps all address arithmetic in 32 bits and stops
// the compiler from re-deriving the carve-up inside the hot loops.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// Shared-memory carve-up, identical for forward and backward.
//   tile [T+1][HS]  staged rows; row T is all zeros (target of padded gather slots -> branch-free inner loop)
//   meta [T]        .x = CSR row start | (row length << 16), .y = bits of 1/denom (0 for unobservable rows)
//   col  [4T+2]     16-bit column indices (3T used; the tail is slack so that padded slots read in bounds)
struct Smem {
    float* tile;
    uint2* meta;
    unsigned short* col;
};
__host__ __device__ inline size_t tile_rows(int T, int lpr) {
    const int groups = kWarps * (32 / lpr);
    return (size_t)(T + 1 > groups ? T + 1 : groups);
}
__host__ __device__ inline size_t agg_smem_bytes(int T, int lpr) {
    return tile_rows(T, lpr) * 4 * lpr * sizeof(float) + (size_t)T * sizeof(uint2) +
           (size_t)(4 * T + 2) * sizeof(unsigned short);
}
__device__ __forceinline__ Smem carve(float* base, int T, int lpr) {
    Smem s;
    s.tile = base;
    s.meta = reinterpret_cast<uint2*>(base + tile_rows(T, lpr) * 4 * lpr);
    s.col = reinterpret_cast<unsigned short*>(s.meta + T);
    return s;
}

// Stage the sentence's CSR (16-bit) and the per-row {start, length, 1/denom} words; zero the pad row.
template <bool FWD>
__device__ __forceinline__ void stage_meta(const AggParams& p, int b, const Smem& s, int HS) {
    const int T = p.T;
    const int* rp = p.rowptr + (size_t)b * (T + 1);
    const int nnz = p.use_adj ? min(rp[T], p.cap) : 0;
    const int* cb = p.col + (size_t)b * p.cap;
    for (int e = threadIdx.x; e < nnz; e += kThreads) s.col[e] = (unsigned short)cb[e];
    for (int t = threadIdx.x; t < T; t += kThreads) {
        const int st = rp[t], len = p.use_adj ? rp[t + 1] - st : 0;
        float inv = 0.f;
        if (FWD) {
            const bool on = p.flags[(size_t)b * T + t] != 0;  // observable row: in the tree, or an entity token
            inv = on ? __frcp_rn(p.denom[(size_t)b * T + t]) : 0.f;
        }
        s.meta[t] = make_uint2((unsigned)st | ((unsigned)len << 16), __float_as_uint(inv));
    }
    for (int c = threadIdx.x; c < HS; c += kThreads) s.tile[(size_t)T * HS + c] = 0.f;
}

// acc0/acc1 += the rows listed in two CSR rows.  The trip count is warp-uniform (max row length in the warp),
// exhausted slots select the zero row, so the loop has no divergent control flow:
//   per chain and trip: LDS.U16, ISETP+SEL, IMAD, LDS.128, 4 FADD.
template <int HS>
__device__ __forceinline__ void gather2(uint32_t tile_lane, uint32_t col_s, int T, unsigned m0, unsigned m1,
                                        float4& acc0, float4& acc1) {
    const int n0 = m0 >> 16, n1 = m1 >> 16;
    uint32_t c0 = col_s + 2u * (m0 & 0xffffu), c1 = col_s + 2u * (m1 & 0xffffu);
    const int trips = __reduce_max_sync(GPT_FULL_MASK, max(n0, n1));
#pragma unroll 2
    for (int k = 0; k < trips; ++k) {
        const uint32_t r0 = lds_u16(c0), r1 = lds_u16(c1);
        const uint32_t j0 = (k < n0) ? r0 : (uint32_t)T, j1 = (k < n1) ? r1 : (uint32_t)T;
        const float4 x0 = lds128(tile_lane + j0 * (HS * 4)), x1 = lds128(tile_lane + j1 * (HS * 4));
        add4(acc0, x0);
        add4(acc1, x1);
        c0 += 2;
        c1 += 2;
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

enum { DROP_NONE = 0, DROP_PHILOX = 1, DROP_MASK = 2 };

template <int LPR, bool ALIGNED, int DROP>
__global__ void __launch_bounds__(kThreads) aggregate_fwd_kernel(const AggParams p) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int HS = 4 * LPR, RPW = 32 / LPR, GROUPS = kWarps * RPW;
    const int T = p.T, H = p.H;
    const int b = blockIdx.y, col0 = blockIdx.x * HS;
    const Smem s = carve(smem_f, T, LPR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage y[b, :, col0 : col0+HS] -----------------------------------------------------------------------
    const float* yb = p.y + (size_t)b * T * H;
    if (ALIGNED) {
        for (int q = tid; q < T * LPR; q += kThreads) {
            const int row = q / LPR, cc = (q % LPR) * 4, c = col0 + cc;
            float* dst = s.tile + row * HS + cc;
            if (c < H) cp_async16(dst, yb + (size_t)row * H + c);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int q = tid; q < T * HS; q += kThreads) {
            const int row = q / HS, cc = q % HS, c = col0 + cc;
            if (c < H) cp_async4(s.tile + q, yb + (size_t)row * H + c);
            else s.tile[q] = 0.f;
        }
    }
    stage_meta<true>(p, b, s, HS);
    cp_async_wait_all();
    __syncthreads();

    // ---- per-thread constants ---------------------------------------------------------------------------------
    const int cl = (lane % LPR) * 4, c_lane = col0 + cl, sub = lane / LPR;
    const bool col_ok = c_lane < H;  // lanes past H stay in the loops (warp-wide reduce inside) but never store
    float4 bias2;
    bias2.x = col_ok ? 2.0f * p.bias[c_lane] : 0.f;
    bias2.y = (c_lane + 1 < H) ? 2.0f * p.bias[c_lane + 1] : 0.f;
    bias2.z = (c_lane + 2 < H) ? 2.0f * p.bias[c_lane + 2] : 0.f;
    bias2.w = (c_lane + 3 < H) ? 2.0f * p.bias[c_lane + 3] : 0.f;
    unsigned long long seed = 0, step = 0;
    if (DROP == DROP_PHILOX) { seed = p.rng[0]; step = p.rng[1]; }
    const uint32_t tile_lane = smem_u32(s.tile) + (uint32_t)cl * 4u;
    const uint32_t meta_s = smem_u32(s.meta), col_s = smem_u32(s.col);
    const uint32_t zero_row = tile_lane + (uint32_t)T * (HS * 4);
    float* const out_lane = p.out + (size_t)b * T * H + c_lane;
    const float* const mask_lane = (DROP == DROP_MASK) ? p.drop_mask + (size_t)b * T * H + c_lane : nullptr;
    const unsigned thresh = p.thresh16;
    const float dscale = p.drop_scale;

    // ---- LPR lanes per node; every lane group walks its own blocks of kRowBlock consecutive rows.  The block loop
    //      is warp-uniform (blk0); rows past T behave as empty rows and are not stored. ---------------------------
    for (int blk0 = warp * RPW; blk0 * kRowBlock < T; blk0 += GROUPS) {
        const int blk = blk0 + sub;
        unsigned long long rlo[4] = {0, 0, 0, 0}, rhi[4] = {0, 0, 0, 0};  // 8 x 16 random bits per column
        if (DROP == DROP_PHILOX) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const Philox4 q = philox4x32(  // one call covers this column for the 8 rows of the block
                    (uint32_t)(c_lane + v) | (p.subseq << 20), (uint32_t)blk, (uint32_t)b, (uint32_t)step,
                    (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
                rlo[v] = (unsigned long long)q.x | ((unsigned long long)q.y << 32);
                rhi[v] = (unsigned long long)q.z | ((unsigned long long)q.w << 32);
            }
        }
#pragma unroll
        for (int r = 0; r < kRowBlock; r += 2) {
            const int i0 = blk * kRowBlock + r, i1 = i0 + 1;
            const bool v0 = i0 < T, v1 = i1 < T;
            uint2 m0 = make_uint2(0u, 0u), m1 = make_uint2(0u, 0u);
            if (v0) m0 = lds64(meta_s + (uint32_t)i0 * 8u);
            if (v1) m1 = lds64(meta_s + (uint32_t)i1 * 8u);
            // the separate W(h) self term (the CSR row holds the 84-diagonal a second time)
            float4 acc0 = lds128(v0 ? tile_lane + (uint32_t)i0 * (HS * 4) : zero_row);
            float4 acc1 = lds128(v1 ? tile_lane + (uint32_t)i1 * (HS * 4) : zero_row);
            gather2<HS>(tile_lane, col_s, T, m0.x, m1.x, acc0, acc1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 a = h ? acc1 : acc0;
                const float inv = __uint_as_float(h ? m1.y : m0.y);  // 0 for unobservable rows -> output 0
                float res[4] = {(a.x + bias2.x) * inv, (a.y + bias2.y) * inv, (a.z + bias2.z) * inv,
                                (a.w + bias2.w) * inv};
                const int rr = r + h;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float t = fmaxf(res[v], 0.f);
                    if (DROP == DROP_PHILOX) {
                        const uint32_t bits = (uint32_t)(((rr < 4 ? rlo[v] : rhi[v]) >> ((rr & 3) * 16)) & 0xffffull);
                        t = (bits >= thresh) ? t * dscale : 0.f;
                    }
                    res[v] = t;
                }
                if ((h ? v1 : v0) && col_ok) {
                    const uint32_t off = (uint32_t)(h ? i1 : i0) * (uint32_t)H;
                    if (DROP == DROP_MASK) {
                        const float* m = mask_lane + off;
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            if (c_lane + v < H) res[v] *= m[v];
                    }
                    store4<ALIGNED>(out_lane + off, make_float4(res[0], res[1], res[2], res[3]), c_lane, H);
                }
            }
        }
    }
}

template <int LPR, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) aggregate_bwd_kernel(const AggParams p) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int HS = 4 * LPR, RPW = 32 / LPR, GROUPS = kWarps * RPW;
    const int T = p.T, H = p.H;
    const int b = blockIdx.y, col0 = blockIdx.x * HS;
    const Smem s = carve(smem_f, T, LPR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage g = gout * dropscale * [out > 0] / denom --------------------------------------------------------
    const size_t base = (size_t)b * T * H;
    const float* gb = p.y + base;
    const float* ob = p.aux + base;
    const float* mb = p.drop_mask ? p.drop_mask + base : nullptr;
    const float* dnb = p.denom + (size_t)b * T;
    if (ALIGNED) {
        const uint32_t tile_s = smem_u32(s.tile);
        const float ds = p.drop_scale;
#pragma unroll 4
        for (int q = tid; q < T * LPR; q += kThreads) {
            const int row = q / LPR, cc = (q % LPR) * 4, c = col0 + cc;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < H) {
                const uint32_t off = (uint32_t)row * (uint32_t)H + (uint32_t)c;
                const float4 go = *reinterpret_cast<const float4*>(gb + off);
                const float4 o = *reinterpret_cast<const float4*>(ob + off);
                float4 m = make_float4(ds, ds, ds, ds);
                if (mb) m = *reinterpret_cast<const float4*>(mb + off);
                const float inv = __frcp_rn(dnb[row]);
                g.x = o.x > 0.f ? go.x * m.x * inv : 0.f;
                g.y = o.y > 0.f ? go.y * m.y * inv : 0.f;
                g.z = o.z > 0.f ? go.z * m.z * inv : 0.f;
                g.w = o.w > 0.f ? go.w * m.w * inv : 0.f;
            }
            sts128(tile_s + (uint32_t)(row * HS + cc) * 4u, g);
        }
    } else {
        for (int q = tid; q < T * HS; q += kThreads) {
            const int row = q / HS, cc = q % HS, c = col0 + cc;
            float g = 0.f;
            if (c < H) {
                const size_t off = (size_t)row * H + c;
                const float m = mb ? mb[off] : p.drop_scale;
                g = ob[off] > 0.f ? gb[off] * m * __frcp_rn(dnb[row]) : 0.f;
            }
            s.tile[q] = g;
        }
    }
    stage_meta<false>(p, b, s, HS);
    __syncthreads();

    // ---- dy_j = g_j + sum_{i in row j} g_i ; column sums for dbias ---------------------------------------------
    const int cl = (lane % LPR) * 4, c_lane = col0 + cl, sub = lane / LPR;
    const bool col_ok = c_lane < H;
    const uint32_t tile_lane = smem_u32(s.tile) + (uint32_t)cl * 4u;
    const uint32_t meta_s = smem_u32(s.meta), col_s = smem_u32(s.col);
    const uint32_t zero_row = tile_lane + (uint32_t)T * (HS * 4);
    float* const out_lane = p.out + base + c_lane;
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jb = warp * RPW * 2; jb < T; jb += GROUPS * 2) {   // warp-uniform
        const int j0 = jb + sub * 2, j1 = j0 + 1;
        const bool v0 = j0 < T, v1 = j1 < T;
        unsigned m0 = 0u, m1 = 0u;
        if (v0) m0 = lds64(meta_s + (uint32_t)j0 * 8u).x;
        if (v1) m1 = lds64(meta_s + (uint32_t)j1 * 8u).x;
        float4 acc0 = lds128(v0 ? tile_lane + (uint32_t)j0 * (HS * 4) : zero_row);
        float4 acc1 = lds128(v1 ? tile_lane + (uint32_t)j1 * (HS * 4) : zero_row);
        add4(csum, acc0);
        add4(csum, acc1);
        gather2<HS>(tile_lane, col_s, T, m0, m1, acc0, acc1);
        const uint32_t off = (uint32_t)j0 * (uint32_t)H;
        if (v0 && col_ok) store4<ALIGNED>(out_lane + off, acc0, c_lane, H);
        if (v1 && col_ok) store4<ALIGNED>(out_lane + off + H, acc1, c_lane, H);
    }
    if (p.dbias != nullptr) {
        __syncthreads();  // everyone is done reading the tile; reuse its head for the cross-group reduction
        float* red = s.tile;  // [GROUPS][HS]
        *reinterpret_cast<float4*>(red + (warp * RPW + sub) * HS + cl) = csum;
        __syncthreads();
        for (int c = tid; c < HS; c += kThreads) {
            float sum = 0.f;
#pragma unroll 8
            for (int g = 0; g < GROUPS; ++g) sum += red[g * HS + c];
            if (col0 + c < H) atomicAdd(p.dbias + col0 + c, 2.0f * sum);  // the bias enters the layer twice
        }
    }
}

// widest slice that still leaves >= 3 CTAs per SM and fills the machine at least ~2 waves
int pick_lpr(int B, int T, int H, int force_vec) {
    if (force_vec == 1) return 8;
    if (force_vec == 2) return 16;
    if (force_vec == 4) return 32;
    for (int lpr = 32; lpr >= 8; lpr >>= 1) {
        if (agg_smem_bytes(T, lpr) > 74 * 1024) continue;  // 3 CTAs per SM
        const long slices = (H + 4 * lpr - 1) / (4 * lpr);
        if (lpr > 8 && slices * B < 2 * 148) continue;
        return lpr;
    }
    return 8;
}

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t a = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (a != cudaSuccess) return (int)a;
    }
    return GPT_OK;
}

template <int LPR, bool ALIGNED>
int launch(bool fwd, const AggParams& p, cudaStream_t st) {
    const size_t smem = agg_smem_bytes(p.T, LPR);
    if (smem > 224 * 1024) return GPT_ERR_UNSUPPORTED;
    dim3 grid((p.H + 4 * LPR - 1) / (4 * LPR), p.B);
    int rc;
    if (fwd) {
        if (p.drop_mask != nullptr) {
            if ((rc = ensure_smem(aggregate_fwd_kernel<LPR, ALIGNED, DROP_MASK>, smem)) != GPT_OK) return rc;
            aggregate_fwd_kernel<LPR, ALIGNED, DROP_MASK><<<grid, kThreads, smem, st>>>(p);
        } else if (p.rng != nullptr && p.thresh16 > 0) {
            if ((rc = ensure_smem(aggregate_fwd_kernel<LPR, ALIGNED, DROP_PHILOX>, smem)) != GPT_OK) return rc;
            aggregate_fwd_kernel<LPR, ALIGNED, DROP_PHILOX><<<grid, kThreads, smem, st>>>(p);
        } else {
            if ((rc = ensure_smem(aggregate_fwd_kernel<LPR, ALIGNED, DROP_NONE>, smem)) != GPT_OK) return rc;
            aggregate_fwd_kernel<LPR, ALIGNED, DROP_NONE><<<grid, kThreads, smem, st>>>(p);
        }
    } else {
        if ((rc = ensure_smem(aggregate_bwd_kernel<LPR, ALIGNED>, smem)) != GPT_OK) return rc;
        aggregate_bwd_kernel<LPR, ALIGNED><<<grid, kThreads, smem, st>>>(p);
    }
    return gpt_launch_status();
}

int dispatch(bool fwd, const AggParams& p, int force_vec, cudaStream_t st) {
    if (4 * p.T + 2 > 65535) return GPT_ERR_UNSUPPORTED;  // 16-bit CSR indices / row lengths in shared memory
    const bool aligned = (p.H % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                         (p.aux == nullptr || (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) &&
                         (p.drop_mask == nullptr || (reinterpret_cast<uintptr_t>(p.drop_mask) & 15) == 0);
    int lpr = pick_lpr(p.B, p.T, p.H, force_vec);
    while (lpr > 8 && agg_smem_bytes(p.T, lpr) > 224 * 1024) lpr >>= 1;
    if (aligned) {
        if (lpr == 32) return launch<32, true>(fwd, p, st);
        if (lpr == 16) return launch<16, true>(fwd, p, st);
        return launch<8, true>(fwd, p, st);
    }
    if (lpr == 32) return launch<32, false>(fwd, p, st);
    if (lpr == 16) return launch<16, false>(fwd, p, st);
    return launch<8, false>(fwd, p, st);
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
