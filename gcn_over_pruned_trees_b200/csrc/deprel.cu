// K10 -- relation-aware GCN layers over the pruned-tree CSR: adj_type 'full_deprel' and 'diagonal_deprel'.
//
// Replaces, per layer, the reference's dense path (all of it on [B,T,T] float matrices and [B,T,D,K] outer products)
//   full_deprel      /root/reference/model/gcn.py:296-386, traverse_deprel :400-415, traverse_self_loop :417-434,
//                    maybe_drop_edges :436-449, maybe_forget_deprels :451-470
//   diagonal_deprel  /root/reference/model/gcn.py:272-294
//   AxW / denom ; relu ; gcn_drop   /root/reference/model/gcn.py:390-393
//
// Restated.  With Z[n,d,:] = x_n . weight_l[d] (ONE projection GEMM per layer, K3, shared by the
// three directions -- the reference runs the [B,T,D,K] x [D,K,H] contraction three times) and e(.) the relation
// vectors,
//   F_n = sum_d e(deprel_n)[d]      (Z[n,d,:] + bias_l[d])     "forward":  what a PARENT collects from child n
//   R_n = sum_d e(deprel_n + 42)[d] (Z[n,d,:] + bias_l[d])     "reverse":  what a CHILD collects from its parent n --
//                                                               keyed by the parent's own relation (gcn.py:349)
//   S_n = sum_d e(84)[d]            (Z[n,d,:] + bias_l[d])     self loop
//   out_i = dropout(relu((sum_{c child of i} keep(i,c) F_c + keep(i,p) R_{p = parent of i} + S_i) / denom_i))
// (diagonal_deprel: F_n = e(deprel_n) * x_n etc., elementwise, no edge dropout.)  The CSR entry's value (K1's `val`)
// tells the direction: 0 < val < 42 parent->child, 42 < val < 84 child->parent (tree.py:184-192).  For layers
// l >= deprel_max_depth, and for tokens whose relation is "forgotten", the relation vector is all ones.
// Rows with flags == 0 (outside the pruned tree and not an entity token) are written as zeros: they are not neighbours
// of kept tokens and are masked out of all three pools, so they never reach the logits (same convention as K2).  An
// entity token outside the tree (singleton tree; entity outside the last root's component for prune_k < 0) has an
// empty CSR row and is computed as the reference computes it -- relu(S_i / 1) -- because the subject / object pools
// read it.
//
// Mapping: one CTA per token row (grid-stride), threads over the H output columns, every global access coalesced
// along H.  HBM-bound integer/fp32 work: relmix streams Z once ([N, D*H], the dominant traffic), agg3 reads each
// F/R/S row once per incident edge out of L2.
//
// Backward (symmetric CSR: a row lists the parent and the children of its token, which is all a gather needs):
//   g_i  = gout_i * d(out_i)/d(z_i) / denom_i
//   dS_j = g_j ,  dF_j = keep(p,j) g_p (p = parent of j) ,  dR_j = sum_{c child of j} keep(c,j) g_c
//   dZ[n,d,:] = e_f[d] dF_n + e_r[d] dR_n + e_s[d] dS_n ,  de(.)[d] += <d{F,R,S}_n, Z[n,d,:] + bias_l[d]>
#include "gpt_common.cuh"
#include <cstdlib>

namespace {

constexpr int kThreads = 128;
constexpr int kFwdBound = 42;   // /root/reference/utils/constant.py:14
constexpr int kRevBound = 84;   // constant.py:16; also the self-loop relation id (constant.py:17)

__device__ __forceinline__ int rel_id(long long r) { return (r < 0 || r > kFwdBound) ? 0 : (int)r; }
// a row some pool can see: in the tree, or a subject / object token (flags of gpt_prune_csr)
__device__ __forceinline__ bool observable(unsigned char f) { return f != 0; }

// Bernoulli(keep_prob) of entry [b, i, j] of the dense matrix of direction `dir` (0: parent->child matrix, 1: child->parent)
// at layer `layer`: explicit dense masks when given (tests), else Philox keyed by {seed, step}.
__device__ __forceinline__ bool edge_kept(const unsigned char* __restrict__ dense, const unsigned long long* __restrict__ rng,
                                          float keep_prob, unsigned layer, unsigned dir, int b, int i, int j, int T) {
    if (dense != nullptr) return dense[((size_t)b * T + i) * T + j] != 0;
    if (keep_prob >= 1.f || rng == nullptr) return true;
    const unsigned long long seed = rng[0], step = rng[1];
    const Philox4 r = philox4x32((uint32_t)(b * T + i), (uint32_t)j, 0xED6E0000u | (layer << 1) | dir, (uint32_t)step,
                                 (uint32_t)seed, (uint32_t)(seed >> 32));
    return u01(r.x) < keep_prob;
}

// ---- relation mix: Z -> F, R, S ------------------------------------------------------------------------------------
struct MixParams {
    const float* Z;               // [N, D*H] projected rows (no bias)
    const float* bias;            // [D*H]
    const float* E;               // [85, D] relation vectors
    const long long* deprel;      // [N]
    const unsigned char* flags;   // [N]
    const unsigned char* keep_f;  // optional [N]: 0 = this token's forward relation vector is forgotten (all ones)
    const unsigned char* keep_r;  // optional [N]: same for the reverse direction
    float* F;                     // fwd: outputs [N,H]; bwd: dF, dR, dS
    float* R;
    float* S;
    float* dZ;                    // bwd: [N, D*H]
    float* dE;                    // bwd: [85, D], caller-zeroed, atomically accumulated
    const int* perm;              // optional: the observable rows, compacted (gpt_live_rows) -- Z / dZ are then indexed by
    const int* count;             //   the compact position i < *count, everything per token by n = perm[i]
    int N, D, H, deep;
};

// rows this launch walks: every token row, or only the compacted observable ones
__device__ __forceinline__ int mix_rows(const MixParams& p) { return p.perm != nullptr ? min(*p.count, p.N) : p.N; }

__device__ __forceinline__ void load_relation_vectors(const MixParams& p, int n, int rf, float* ef, float* er, float* es) {
    const bool forget_f = p.deep || (p.keep_f != nullptr && p.keep_f[n] == 0);
    const bool forget_r = p.deep || (p.keep_r != nullptr && p.keep_r[n] == 0);
    for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
        ef[d] = forget_f ? 1.f : p.E[(size_t)rf * p.D + d];
        er[d] = forget_r ? 1.f : p.E[(size_t)(rf + kFwdBound) * p.D + d];
        es[d] = p.deep ? 1.f : p.E[(size_t)kRevBound * p.D + d];
    }
}

__global__ void __launch_bounds__(kThreads) relmix_fwd_kernel(const MixParams p) {
    extern __shared__ float sm[];
    float* ef = sm;
    float* er = sm + p.D;
    float* es = sm + 2 * p.D;
    GPT_PDL_ENTER();
    const size_t DH = (size_t)p.D * p.H;
    const int rows = mix_rows(p);
    for (int i = blockIdx.x; i < rows; i += gridDim.x) {
        const int n = p.perm != nullptr ? p.perm[i] : i;
        if (!observable(p.flags[n])) {                 // CTA-uniform
            for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
                p.F[(size_t)n * p.H + h] = 0.f;
                p.R[(size_t)n * p.H + h] = 0.f;
                p.S[(size_t)n * p.H + h] = 0.f;
            }
            continue;
        }
        load_relation_vectors(p, n, rel_id(p.deprel[n]), ef, er, es);
        __syncthreads();
        for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
            const float* __restrict__ z = p.Z + (size_t)i * DH + h;
            float f = 0.f, r = 0.f, s = 0.f;
            for (int d = 0; d < p.D; ++d) {
                const float v = z[(size_t)d * p.H] + (p.bias != nullptr ? p.bias[(size_t)d * p.H + h] : 0.f);
                f = fmaf(ef[d], v, f);
                r = fmaf(er[d], v, r);
                s = fmaf(es[d], v, s);
            }
            p.F[(size_t)n * p.H + h] = f;
            p.R[(size_t)n * p.H + h] = r;
            p.S[(size_t)n * p.H + h] = s;
        }
        __syncthreads();                                       // ef/er/es are rewritten for the next row
    }
}

// smem: ef, er, es [D] | dF, dR, dS [H] | acc84 [D]
__global__ void __launch_bounds__(kThreads) relmix_bwd_kernel(const MixParams p) {
    extern __shared__ float sm[];
    float* ef = sm;
    float* er = sm + p.D;
    float* es = sm + 2 * p.D;
    float* gf = sm + 3 * p.D;
    float* gr = gf + p.H;
    float* gs = gr + p.H;
    float* acc84 = gs + p.H;      // this CTA's share of the self-loop vector's gradient, flushed once at the end
    GPT_PDL_ENTER();
    const size_t DH = (size_t)p.D * p.H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int d = threadIdx.x; d < p.D; d += blockDim.x) acc84[d] = 0.f;
    __syncthreads();
    const int rows = mix_rows(p);
    for (int i = blockIdx.x; i < rows; i += gridDim.x) {
        const int n = p.perm != nullptr ? p.perm[i] : i;
        float* __restrict__ dz = p.dZ + (size_t)i * DH;
        if (!observable(p.flags[n])) {                 // CTA-uniform; the projection's gradients read every row
            for (size_t i = threadIdx.x; i < DH; i += blockDim.x) dz[i] = 0.f;
            continue;
        }
        const int rf = rel_id(p.deprel[n]);
        const bool forget_f = p.deep || (p.keep_f != nullptr && p.keep_f[n] == 0);
        const bool forget_r = p.deep || (p.keep_r != nullptr && p.keep_r[n] == 0);
        load_relation_vectors(p, n, rf, ef, er, es);
        for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
            gf[h] = p.F[(size_t)n * p.H + h];
            gr[h] = p.R[(size_t)n * p.H + h];
            gs[h] = p.S[(size_t)n * p.H + h];
        }
        __syncthreads();
        const float* __restrict__ z = p.Z + (size_t)i * DH;
        for (int d = warp; d < p.D; d += nwarps) {             // a relation slot d always belongs to the same warp
            const float a = ef[d], b = er[d], c = es[d];
            float sf = 0.f, sr = 0.f, ss = 0.f;
            for (int h = lane; h < p.H; h += 32) {
                const float v = z[(size_t)d * p.H + h] + (p.bias != nullptr ? p.bias[(size_t)d * p.H + h] : 0.f);
                const float x = gf[h], y = gr[h], w = gs[h];
                sf = fmaf(x, v, sf);
                sr = fmaf(y, v, sr);
                ss = fmaf(w, v, ss);
                dz[(size_t)d * p.H + h] = fmaf(a, x, fmaf(b, y, c * w));
            }
            sf = warp_sum_f(sf);
            sr = warp_sum_f(sr);
            ss = warp_sum_f(ss);
            if (lane == 0 && !p.deep) {
                if (!forget_f && rf != 0) atomicAdd(p.dE + (size_t)rf * p.D + d, sf);      // row 0 is padding_idx
                if (!forget_r) atomicAdd(p.dE + (size_t)(rf + kFwdBound) * p.D + d, sr);
                acc84[d] += ss;
            }
        }
        __syncthreads();
    }
    if (!p.deep)
        for (int d = warp; d < p.D; d += nwarps)
            if (lane == 0 && acc84[d] != 0.f) atomicAdd(p.dE + (size_t)kRevBound * p.D + d, acc84[d]);
}

// ---- relation mix, 128-bit path (H % 4 == 0, 16-byte aligned rows): warps over the relation slots, lanes over the columns ----
constexpr int kMixThreads = 256;
__device__ __forceinline__ float4 f4_fma(float a, float4 v, float4 acc) {
    return make_float4(fmaf(a, v.x, acc.x), fmaf(a, v.y, acc.y), fmaf(a, v.z, acc.z), fmaf(a, v.w, acc.w));
}
// volatile: ptxas keeps these in program order, so a batch of loads is in flight before its first consumer
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
#ifdef GPT_HOST_EMULATION
    return *p;
#else
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b, float acc) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}

// smem: ef, er, es [D] | part [warps][3][H]: warp w sums the slots d = w, w + warps, ...; the partial rows meet in shared memory
__global__ void __launch_bounds__(kMixThreads) relmix_fwd4_kernel(const MixParams p) {
    extern __shared__ float sm[];
    float* ef = sm;
    float* er = sm + p.D;
    float* es = sm + 2 * p.D;
    float4* part = reinterpret_cast<float4*>(sm + ((3 * p.D + 3) & ~3));
    GPT_PDL_ENTER();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5, hq = p.H >> 2;
    const size_t DH = (size_t)p.D * p.H;
    const float4* __restrict__ bias4 = reinterpret_cast<const float4*>(p.bias);
    const bool has_b = p.bias != nullptr;       // else the projection's epilogue already added it to Z
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const int rows = mix_rows(p);
    for (int i = blockIdx.x; i < rows; i += gridDim.x) {
        const int n = p.perm != nullptr ? p.perm[i] : i;
        float4* __restrict__ F4 = reinterpret_cast<float4*>(p.F + (size_t)n * p.H);
        float4* __restrict__ R4 = reinterpret_cast<float4*>(p.R + (size_t)n * p.H);
        float4* __restrict__ S4 = reinterpret_cast<float4*>(p.S + (size_t)n * p.H);
        if (!observable(p.flags[n])) {                 // CTA-uniform
            for (int c = threadIdx.x; c < hq; c += blockDim.x) F4[c] = R4[c] = S4[c] = zero;
            continue;
        }
        load_relation_vectors(p, n, rel_id(p.deprel[n]), ef, er, es);
        __syncthreads();
        const float4* __restrict__ z4 = reinterpret_cast<const float4*>(p.Z + (size_t)i * DH);
        for (int c = lane; c < hq; c += 32) {
            float4 f = zero, r = zero, s = zero;
#pragma unroll 4
            for (int d = warp; d < p.D; d += nw) {
                const float4 z = z4[(size_t)d * hq + c], b = has_b ? bias4[(size_t)d * hq + c] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 v = make_float4(z.x + b.x, z.y + b.y, z.z + b.z, z.w + b.w);
                f = f4_fma(ef[d], v, f);
                r = f4_fma(er[d], v, r);
                s = f4_fma(es[d], v, s);
            }
            part[(size_t)(warp * 3 + 0) * hq + c] = f;
            part[(size_t)(warp * 3 + 1) * hq + c] = r;
            part[(size_t)(warp * 3 + 2) * hq + c] = s;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 3 * hq; idx += blockDim.x) {
            const int which = idx / hq, c = idx - which * hq;
            float4 acc = zero;
            for (int w = 0; w < nw; ++w) {
                const float4 v = part[(size_t)(w * 3 + which) * hq + c];
                acc = make_float4(acc.x + v.x, acc.y + v.y, acc.z + v.z, acc.w + v.w);
            }
            (which == 0 ? F4 : which == 1 ? R4 : S4)[c] = acc;
        }
        __syncthreads();                                       // ef/er/es and part are rewritten for the next row
    }
}

// smem: ef, er, es [D] | acc84 [D] | dF, dR, dS [H]
__global__ void __launch_bounds__(kMixThreads, 3) relmix_bwd4_kernel(const MixParams p) {
    extern __shared__ float sm[];
    float* ef = sm;
    float* er = sm + p.D;
    float* es = sm + 2 * p.D;
    float* acc84 = sm + 3 * p.D;
    float4* gf = reinterpret_cast<float4*>(sm + ((4 * p.D + 3) & ~3));
    const int hq = p.H >> 2;
    float4* gr = gf + hq;
    float4* gs = gr + hq;
    GPT_PDL_ENTER();
    const size_t DH = (size_t)p.D * p.H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float4* __restrict__ bias4 = reinterpret_cast<const float4*>(p.bias);
    const bool has_b = p.bias != nullptr;       // else the projection's epilogue already added it to Z
    for (int d = threadIdx.x; d < p.D; d += blockDim.x) acc84[d] = 0.f;
    __syncthreads();
    const int rows = mix_rows(p);
    for (int i = blockIdx.x; i < rows; i += gridDim.x) {
        const int n = p.perm != nullptr ? p.perm[i] : i;
        float4* __restrict__ dz4 = reinterpret_cast<float4*>(p.dZ + (size_t)i * DH);
        if (!observable(p.flags[n])) {                 // CTA-uniform; the projection's gradients read every row
            for (size_t k = threadIdx.x; k < DH / 4; k += blockDim.x) dz4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const int rf = rel_id(p.deprel[n]);
        const bool forget_f = p.deep || (p.keep_f != nullptr && p.keep_f[n] == 0);
        const bool forget_r = p.deep || (p.keep_r != nullptr && p.keep_r[n] == 0);
        load_relation_vectors(p, n, rf, ef, er, es);
        for (int c = threadIdx.x; c < hq; c += blockDim.x) {
            gf[c] = reinterpret_cast<const float4*>(p.F + (size_t)n * p.H)[c];
            gr[c] = reinterpret_cast<const float4*>(p.R + (size_t)n * p.H)[c];
            gs[c] = reinterpret_cast<const float4*>(p.S + (size_t)n * p.H)[c];
        }
        __syncthreads();
        const float4* __restrict__ z4 = reinterpret_cast<const float4*>(p.Z + (size_t)i * DH);
        // U relation slots x 2 column groups per lane: every load of Z is issued before the first store of dZ (the stores
        // would otherwise fence the loads behind them -- one round trip to memory per slot)
        constexpr int U = 4;
        for (int db = warp; db < p.D; db += nw * U) {          // a relation slot d always belongs to the same warp
            float sf[U], sr[U], ss[U];
#pragma unroll
            for (int u = 0; u < U; ++u) sf[u] = sr[u] = ss[u] = 0.f;
            for (int q0 = lane; q0 < hq; q0 += 64) {
                float4 zv[U][2];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int d = db + u * nw, q = q0 + 32 * j;
                        if (d < p.D && q < hq) zv[u][j] = ld_stream_f4(z4 + (size_t)d * hq + q);
                    }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int d = db + u * nw;
                    if (d >= p.D) break;                       // warp-uniform
                    const float a = ef[d], b = er[d], c = es[d];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int q = q0 + 32 * j;
                        if (q >= hq) break;
                        const float4 z = zv[u][j], bb = has_b ? bias4[(size_t)d * hq + q] : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 v = make_float4(z.x + bb.x, z.y + bb.y, z.z + bb.z, z.w + bb.w);
                        const float4 x = gf[q], y = gr[q], w = gs[q];
                        sf[u] = f4_dot(x, v, sf[u]);
                        sr[u] = f4_dot(y, v, sr[u]);
                        ss[u] = f4_dot(w, v, ss[u]);
                        dz4[(size_t)d * hq + q] =
                            make_float4(fmaf(a, x.x, fmaf(b, y.x, c * w.x)), fmaf(a, x.y, fmaf(b, y.y, c * w.y)),
                                        fmaf(a, x.z, fmaf(b, y.z, c * w.z)), fmaf(a, x.w, fmaf(b, y.w, c * w.w)));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int d = db + u * nw;
                if (d >= p.D) break;
                const float tf = warp_sum_f(sf[u]), tr = warp_sum_f(sr[u]), ts = warp_sum_f(ss[u]);
                if (lane == 0 && !p.deep) {
                    if (!forget_f && rf != 0) atomicAdd(p.dE + (size_t)rf * p.D + d, tf);      // row 0 is padding_idx
                    if (!forget_r) atomicAdd(p.dE + (size_t)(rf + kFwdBound) * p.D + d, tr);
                    acc84[d] += ts;
                }
            }
        }
        __syncthreads();
    }
    if (!p.deep)
        for (int d = warp; d < p.D; d += nw)
            if (lane == 0 && acc84[d] != 0.f) atomicAdd(p.dE + (size_t)kRevBound * p.D + d, acc84[d]);
}

// ---- relation mix over the compacted rows: rows streamed through shared memory by bulk copies ---------------------------
// One row is 40 KB of Z (D = 50, H = 200) behind a chain of dependent loads: perm[i] -> deprel[n], keep[n] -> E rows (and,
// backward, the dF / dR / dS rows).  With one row per CTA and per-thread loads that chain and the load round trips, not
// the 40 KB, are the row's time (ncu: every sample on a long-scoreboard stall, 10 % of DRAM bandwidth).  Here a CTA walks
// rows i, i + grid, ... and a NINTH warp runs one row ahead of the eight compute warps: it fetches the next row's vectors
// into the other half of a double buffer and brings the next row of Z in with ONE bulk copy (cp.async.bulk, completion on
// an mbarrier) -- 2 x 40 KB per CTA and two CTAs per SM keep ~160 KB per SM in flight without a register being spent on it.
constexpr int kPipeThreads = kMixThreads + 32;

struct RowMeta {
    int n, rf, forget_f, forget_r;
};

#ifdef GPT_HOST_EMULATION
// host build (tests/emu): the staging copy is done by the fetch warp's lanes, synchronously; the row barrier orders it
__device__ inline void stage_init(unsigned long long*, int) {}
__device__ inline void stage_expect(unsigned long long*, int, int) {}
__device__ inline void stage_row(float* dst, const float* src, int n_floats, unsigned long long*, int lane) {
    for (int k = lane; k < n_floats; k += 32) dst[k] = src[k];
}
__device__ inline void stage_wait(unsigned long long*, unsigned) {}
#else
__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_init(unsigned long long* bars, int lane) {
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_addr(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_addr(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
// one elected lane: expect `total_floats` on the barrier (the sum over every copy of this phase) ...
__device__ __forceinline__ void stage_expect(unsigned long long* bar, int total_floats, int lane) {
    if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_addr(bar)), "r"((uint32_t)total_floats * 4u)
                     : "memory");
}
// ... and one bulk copy global -> shared per call (16-byte aligned, a multiple of 16 bytes)
__device__ __forceinline__ void stage_row(float* dst, const float* src, int n_floats, unsigned long long* bar, int lane) {
    if (lane == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sm_addr(dst)), "l"(src), "r"((uint32_t)n_floats * 4u), "r"(sm_addr(bar)) : "memory");
}
__device__ __forceinline__ void stage_wait(unsigned long long* bar, unsigned parity) {
    uint32_t ok, spins = 0;
    const uint32_t b = sm_addr(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(b), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();      // a lost copy must fail the launch, never hang the GPU
    } while (!ok);
}
#endif

// The fetch warp resolves perm[i] -> deprel[n], keep[n] for 32 of its CTA's rows at once (lane l: the l-th next row), so
// that two of the three dependent round trips of a row's prologue are paid once per 32 rows, not once per row.
__device__ __forceinline__ void pipe_resolve(const MixParams& p, int i0, int rows, RowMeta* table, int lane) {
    const long long i = (long long)i0 + (long long)lane * gridDim.x;
    if (i < rows) {
        const int n = p.perm[i];
        const int rf = rel_id(p.deprel[n]);
        const bool forget_f = p.deep || (p.keep_f != nullptr && p.keep_f[n] == 0);
        const bool forget_r = p.deep || (p.keep_r != nullptr && p.keep_r[n] == 0);
        table[lane] = RowMeta{n, rf, forget_f ? 1 : 0, forget_r ? 1 : 0};
    }
    __syncwarp();
}

// warp `kMixThreads / 32`: the three relation vectors of a resolved row -> buffer.  Two rounds of loads are issued before
// the first store (a store to shared memory orders the loads behind it: one round trip to L2 per round otherwise).
__device__ __forceinline__ void pipe_prefetch(const MixParams& p, const RowMeta m, float* e3, RowMeta* meta, int lane) {
    const float* __restrict__ Ef = p.E + (size_t)m.rf * p.D;
    const float* __restrict__ Er = p.E + (size_t)(m.rf + kFwdBound) * p.D;
    const float* __restrict__ Es = p.E + (size_t)kRevBound * p.D;
    for (int d0 = lane; d0 < p.D; d0 += 64) {
        float v[2][3];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int d = d0 + 32 * j;
            v[j][0] = (d < p.D && !m.forget_f) ? Ef[d] : 1.f;
            v[j][1] = (d < p.D && !m.forget_r) ? Er[d] : 1.f;
            v[j][2] = (d < p.D && !p.deep) ? Es[d] : 1.f;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int d = d0 + 32 * j;
            if (d < p.D) {
                e3[d] = v[j][0];
                e3[p.D + d] = v[j][1];
                e3[2 * p.D + d] = v[j][2];
            }
        }
    }
    if (lane == 0) *meta = m;
}

// backward: the row's incoming gradients dF, dR, dS [H] each -> g3 [3H], three more bulk copies on the same barrier phase
__device__ __forceinline__ void stage_grads(const MixParams& p, int n, float* g3, unsigned long long* bar, int lane) {
    stage_row(g3, p.F + (size_t)n * p.H, p.H, bar, lane);
    stage_row(g3 + p.H, p.R + (size_t)n * p.H, p.H, bar, lane);
    stage_row(g3 + 2 * p.H, p.S + (size_t)n * p.H, p.H, bar, lane);
}

// smem: 2 x Z row [D*H] | 2 x e [3D] | part [8][3][H]
__global__ void __launch_bounds__(kPipeThreads, 2) relmix_fwd_pipe_kernel(const MixParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ RowMeta s_meta[2];
    __shared__ RowMeta s_rows[32];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int e_len = (3 * p.D + 3) & ~3, hq = p.H >> 2, DH = p.D * p.H;
    constexpr int NW = kMixThreads / 32;
    float* z_buf = sm;
    float* e_buf = sm + 2 * DH;
    float4* part = reinterpret_cast<float4*>(e_buf + 2 * e_len);
    GPT_PDL_ENTER();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool fetcher = warp == NW;
    const float4* __restrict__ bias4 = reinterpret_cast<const float4*>(p.bias);
    const bool has_b = p.bias != nullptr;       // else the projection's epilogue already added it to Z
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const int rows = mix_rows(p);
    if ((int)blockIdx.x >= rows) return;
    if (fetcher) {
        stage_init(s_bar, lane);
        stage_expect(&s_bar[0], DH, lane);
        stage_row(z_buf, p.Z + (size_t)blockIdx.x * DH, DH, &s_bar[0], lane);
        pipe_resolve(p, blockIdx.x, rows, s_rows, lane);
        pipe_prefetch(p, s_rows[0], e_buf, &s_meta[0], lane);
    }
    __syncthreads();
    int k = 0;
    for (int i = blockIdx.x; i < rows; i += gridDim.x, ++k) {
        const int cur = k & 1;
        if (fetcher) {
            const int nxt = i + (int)gridDim.x;
            if (nxt < rows) {   // the other buffers were last read before the previous row's barrier
                stage_expect(&s_bar[cur ^ 1], DH, lane);
                stage_row(z_buf + (cur ^ 1) * DH, p.Z + (size_t)nxt * DH, DH, &s_bar[cur ^ 1], lane);
                if (((k + 1) & 31) == 0) pipe_resolve(p, nxt, rows, s_rows, lane);
                pipe_prefetch(p, s_rows[(k + 1) & 31], e_buf + (cur ^ 1) * e_len, &s_meta[cur ^ 1], lane);
            }
        } else {
            const float* ef = e_buf + cur * e_len;
            const float* er = ef + p.D;
            const float* es = er + p.D;
            stage_wait(&s_bar[cur], (unsigned)(k >> 1) & 1u);
            const float4* z4 = reinterpret_cast<const float4*>(z_buf + cur * DH);
            for (int q = lane; q < hq; q += 32) {
                float4 f = zero, r = zero, s = zero;
#pragma unroll 4
                for (int d = warp; d < p.D; d += NW) {
                    const float4 z = z4[d * hq + q], bb = has_b ? bias4[(size_t)d * hq + q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 v = make_float4(z.x + bb.x, z.y + bb.y, z.z + bb.z, z.w + bb.w);
                    f = f4_fma(ef[d], v, f);
                    r = f4_fma(er[d], v, r);
                    s = f4_fma(es[d], v, s);
                }
                part[(warp * 3 + 0) * hq + q] = f;
                part[(warp * 3 + 1) * hq + q] = r;
                part[(warp * 3 + 2) * hq + q] = s;
            }
        }
        __syncthreads();     // the warps' partial rows are complete (and the next row's vectors are in place)
        if (!fetcher) {
            const int n = s_meta[cur].n;
            float4* __restrict__ F4 = reinterpret_cast<float4*>(p.F + (size_t)n * p.H);
            float4* __restrict__ R4 = reinterpret_cast<float4*>(p.R + (size_t)n * p.H);
            float4* __restrict__ S4 = reinterpret_cast<float4*>(p.S + (size_t)n * p.H);
            for (int idx = threadIdx.x; idx < 3 * hq; idx += kMixThreads) {
                const int which = idx / hq, c = idx - which * hq;
                float4 acc = zero;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const float4 v = part[(w * 3 + which) * hq + c];
                    acc = make_float4(acc.x + v.x, acc.y + v.y, acc.z + v.z, acc.w + v.w);
                }
                (which == 0 ? F4 : which == 1 ? R4 : S4)[c] = acc;
            }
        }
        __syncthreads();     // `part` is free again
    }
}

// smem: 2 x Z row [D*H] | acc84 [D] | 2 x { e [3D] | dF, dR, dS [3H] }
__global__ void __launch_bounds__(kPipeThreads, 2) relmix_bwd_pipe_kernel(const MixParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ RowMeta s_meta[2];
    __shared__ RowMeta s_rows[32];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int e_len = (3 * p.D + 3) & ~3, hq = p.H >> 2, acc_len = (p.D + 3) & ~3, DH = p.D * p.H;
    constexpr int NW = kMixThreads / 32;
    float* z_buf = sm;
    float* acc84 = sm + 2 * DH;
    float* bufs = acc84 + acc_len;
    const int buf_len = e_len + 3 * p.H;
    GPT_PDL_ENTER();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool fetcher = warp == NW;
    const float4* __restrict__ bias4 = reinterpret_cast<const float4*>(p.bias);
    const bool has_b = p.bias != nullptr;       // else the projection's epilogue already added it to Z
    const int rows = mix_rows(p);
    if ((int)blockIdx.x >= rows) return;
    for (int d = threadIdx.x; d < p.D; d += blockDim.x) acc84[d] = 0.f;
    if (fetcher) {
        stage_init(s_bar, lane);
        stage_expect(&s_bar[0], DH + 3 * p.H, lane);
        stage_row(z_buf, p.Z + (size_t)blockIdx.x * DH, DH, &s_bar[0], lane);
        pipe_resolve(p, blockIdx.x, rows, s_rows, lane);
        stage_grads(p, s_rows[0].n, bufs + e_len, &s_bar[0], lane);
        pipe_prefetch(p, s_rows[0], bufs, &s_meta[0], lane);
    }
    __syncthreads();
    int k = 0;
    for (int i = blockIdx.x; i < rows; i += gridDim.x, ++k) {
        const int cur = k & 1;
        if (fetcher) {
            const int nxt = i + (int)gridDim.x;
            if (nxt < rows) {
                float* nb = bufs + (cur ^ 1) * buf_len;
                stage_expect(&s_bar[cur ^ 1], DH + 3 * p.H, lane);
                stage_row(z_buf + (cur ^ 1) * DH, p.Z + (size_t)nxt * DH, DH, &s_bar[cur ^ 1], lane);
                if (((k + 1) & 31) == 0) pipe_resolve(p, nxt, rows, s_rows, lane);
                stage_grads(p, s_rows[(k + 1) & 31].n, nb + e_len, &s_bar[cur ^ 1], lane);
                pipe_prefetch(p, s_rows[(k + 1) & 31], nb, &s_meta[cur ^ 1], lane);
            }
        } else {
            const float* ef = bufs + cur * buf_len;
            const float* er = ef + p.D;
            const float* es = er + p.D;
            const float4* gf = reinterpret_cast<const float4*>(ef + e_len);
            const float4* gr = gf + hq;
            const float4* gs = gr + hq;
            const RowMeta m = s_meta[cur];
            stage_wait(&s_bar[cur], (unsigned)(k >> 1) & 1u);
            const float4* z4 = reinterpret_cast<const float4*>(z_buf + cur * DH);
            float4* __restrict__ dz4 = reinterpret_cast<float4*>(p.dZ + (size_t)i * DH);
#pragma unroll 2
            for (int d = warp; d < p.D; d += NW) {             // a relation slot d always belongs to the same warp
                const float a = ef[d], b = er[d], c = es[d];
                float sf = 0.f, sr = 0.f, ss = 0.f;
                for (int q = lane; q < hq; q += 32) {
                    const float4 z = z4[d * hq + q], bb = has_b ? bias4[(size_t)d * hq + q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 v = make_float4(z.x + bb.x, z.y + bb.y, z.z + bb.z, z.w + bb.w);
                    const float4 x = gf[q], y = gr[q], w = gs[q];
                    sf = f4_dot(x, v, sf);
                    sr = f4_dot(y, v, sr);
                    ss = f4_dot(w, v, ss);
                    dz4[(size_t)d * hq + q] =
                        make_float4(fmaf(a, x.x, fmaf(b, y.x, c * w.x)), fmaf(a, x.y, fmaf(b, y.y, c * w.y)),
                                    fmaf(a, x.z, fmaf(b, y.z, c * w.z)), fmaf(a, x.w, fmaf(b, y.w, c * w.w)));
                }
                sf = warp_sum_f(sf);
                sr = warp_sum_f(sr);
                ss = warp_sum_f(ss);
                if (lane == 0 && !p.deep) {
                    if (!m.forget_f && m.rf != 0) atomicAdd(p.dE + (size_t)m.rf * p.D + d, sf);   // row 0 is padding_idx
                    if (!m.forget_r) atomicAdd(p.dE + (size_t)(m.rf + kFwdBound) * p.D + d, sr);
                    acc84[d] += ss;
                }
            }
        }
        __syncthreads();     // this row's buffers are free AND the next row's vectors have landed
    }
    if (!p.deep && !fetcher)
        for (int d = warp; d < p.D; d += NW)
            if (lane == 0 && acc84[d] != 0.f) atomicAdd(p.dE + (size_t)kRevBound * p.D + d, acc84[d]);
}

// ---- the observable rows of a batch, compacted -----------------------------------------------------------------------
// At prune_k = 1 three quarters of a TACRED-shaped batch's token rows are outside every pruned tree: nothing downstream
// reads what the relation-aware layers compute for them, and their [D*H]-wide projections are the step's largest cost.
// One CTA lists the rows with flags != 0 in ascending order:  perm[0 .. count) = their ids, inv[n] = position of row n
// or -1, live[i] = (i < count) -- the K1-style row flags of the COMPACT arrays, which the weight-gradient kernels take --
// and count itself stays on the device: the step is captured into a CUDA graph, so every consumer gets its row count from
// this buffer, never from the host.
__global__ void __launch_bounds__(1024) live_rows_kernel(const unsigned char* __restrict__ flags, int N,
                                                         int* __restrict__ perm, int* __restrict__ inv,
                                                         unsigned char* __restrict__ live, int* __restrict__ count) {
    __shared__ int s_wcnt[32];
    __shared__ int s_base;
    GPT_PDL_ENTER();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < N; base += blockDim.x) {
        const int n = base + tid;
        const bool ob = n < N && observable(flags[n]);
        const unsigned m = __ballot_sync(GPT_FULL_MASK, ob);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < nw; ++w) {
            const int c = s_wcnt[w];
            before += w < warp ? c : 0;
            total += c;
        }
        const int b0 = s_base;
        if (n < N) {
            const int pos = b0 + before + __popc(m & ((1u << lane) - 1u));
            if (ob) perm[pos] = n;
            inv[n] = ob ? pos : -1;
        }
        __syncthreads();
        if (tid == 0) s_base = b0 + total;
        __syncthreads();
    }
    const int total = s_base;
    for (int i = tid; i < N; i += blockDim.x) live[i] = i < total ? 1 : 0;
    if (tid == 0) count[0] = total;
}

// out[i, :] = x[perm[i], :] for i < *count (the other rows of `out` are never read)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ perm,
                                                          const int* __restrict__ count, int N, int K,
                                                          float* __restrict__ out) {
    GPT_PDL_ENTER();
    const int rows = min(*count, N);
    const bool vec = (K % 4 == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    for (int i = blockIdx.x; i < rows; i += gridDim.x) {
        const size_t src = (size_t)perm[i] * K, dst = (size_t)i * K;
        if (vec) {
            for (int k = threadIdx.x; k < K / 4; k += blockDim.x)
                reinterpret_cast<float4*>(out + dst)[k] = reinterpret_cast<const float4*>(x + src)[k];
        } else {
            for (int k = threadIdx.x; k < K; k += blockDim.x) out[dst + k] = x[src + k];
        }
    }
}

// dx[n, :] = inv[n] >= 0 ? dxc[inv[n], :] : 0 for every token row n
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ dxc, const int* __restrict__ inv, int N,
                                                           int K, float* __restrict__ dx) {
    GPT_PDL_ENTER();
    const bool vec = (K % 4 == 0) && (((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dxc)) & 15) == 0);
    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        const int pos = inv[n];
        const size_t src = (size_t)(pos < 0 ? 0 : pos) * K, dst = (size_t)n * K;
        if (vec) {
            for (int k = threadIdx.x; k < K / 4; k += blockDim.x)
                reinterpret_cast<float4*>(dx + dst)[k] =
                    pos < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(dxc + src)[k];
        } else {
            for (int k = threadIdx.x; k < K; k += blockDim.x) dx[dst + k] = pos < 0 ? 0.f : dxc[src + k];
        }
    }
}

// ---- diagonal mode: elementwise relation gates -----------------------------------------------------------------------
struct DiagParams {
    const float* x;               // [N,H]
    const float* E;               // [85,H]
    const long long* deprel;
    const unsigned char* flags;
    float* F;                     // fwd: outputs; bwd: dF, dR, dS
    float* R;
    float* S;
    float* dx;                    // bwd [N,H]
    float* dE;                    // bwd [85,H], caller-zeroed
    int N, H;
};

__global__ void __launch_bounds__(kThreads) diagmix_fwd_kernel(const DiagParams p) {
    GPT_PDL_ENTER();
    for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
        const bool in = observable(p.flags[n]);
        const int rf = rel_id(p.deprel[n]);
        for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
            const size_t o = (size_t)n * p.H + h;
            const float v = in ? p.x[o] : 0.f;
            p.F[o] = v * p.E[(size_t)rf * p.H + h];
            p.R[o] = v * p.E[(size_t)(rf + kFwdBound) * p.H + h];
            p.S[o] = v * p.E[(size_t)kRevBound * p.H + h];
        }
    }
}

// smem: acc84 [H]
__global__ void __launch_bounds__(kThreads) diagmix_bwd_kernel(const DiagParams p) {
    extern __shared__ float sm[];
    GPT_PDL_ENTER();
    for (int h = threadIdx.x; h < p.H; h += blockDim.x) sm[h] = 0.f;   // each thread only ever touches its own h
    for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
        const bool in = observable(p.flags[n]);
        const int rf = rel_id(p.deprel[n]);
        for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
            const size_t o = (size_t)n * p.H + h;
            if (!in) {
                p.dx[o] = 0.f;
                continue;
            }
            const float a = p.F[o], b = p.R[o], c = p.S[o], v = p.x[o];
            p.dx[o] = fmaf(a, p.E[(size_t)rf * p.H + h],
                           fmaf(b, p.E[(size_t)(rf + kFwdBound) * p.H + h], c * p.E[(size_t)kRevBound * p.H + h]));
            if (rf != 0) atomicAdd(p.dE + (size_t)rf * p.H + h, a * v);
            atomicAdd(p.dE + (size_t)(rf + kFwdBound) * p.H + h, b * v);
            sm[h] += c * v;
        }
    }
    for (int h = threadIdx.x; h < p.H; h += blockDim.x)
        if (sm[h] != 0.f) atomicAdd(p.dE + (size_t)kRevBound * p.H + h, sm[h]);
}

// ---- aggregation over the CSR with the layer epilogue -----------------------------------------------------------------
struct Agg3Params {
    const float* F;               // fwd: [N,H] inputs; bwd: unused
    const float* R;
    const float* S;
    const float* gout;            // bwd: d loss / d out [N,H]
    const float* outp;            // bwd: the forward's output (after dropout)
    float* out;                   // fwd: [N,H]
    float* dF;                    // bwd outputs [N,H]
    float* dR;
    float* dS;
    const int* rowptr;            // [B, T+1]
    const int* col;               // [B, cap]
    const unsigned char* val;     // [B, cap]
    const float* denom;           // [N]
    const unsigned char* flags;   // [N]
    const unsigned char* keep_f;  // optional dense [B,T,T] edge masks (tests), one per direction
    const unsigned char* keep_r;
    const unsigned long long* rng;   // {seed, step} on the device: in-kernel edge dropout / dropout
    const float* drop_mask;       // optional explicit, pre-scaled dropout mask [N,H] (tests)
    int B, T, H, cap, directed, self_loop;
    unsigned layer;
    float edge_keep, drop_p;
};

__device__ __forceinline__ bool is_fwd(int v) { return v > 0 && v < kFwdBound; }
__device__ __forceinline__ bool is_rev(int v) { return v > kFwdBound && v < kRevBound; }

__global__ void __launch_bounds__(kThreads) agg3_fwd_kernel(const Agg3Params p) {
    GPT_PDL_ENTER();
    const int N = p.B * p.T;
    const bool philox_drop = p.drop_mask == nullptr && p.drop_p > 0.f && p.rng != nullptr;
    const float scale = philox_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        const int b = n / p.T, i = n - b * p.T;
        if (!observable(p.flags[n])) {
            for (int h = threadIdx.x; h < p.H; h += blockDim.x) p.out[(size_t)n * p.H + h] = 0.f;
            continue;
        }
        const int e0 = p.rowptr[(size_t)b * (p.T + 1) + i], e1 = p.rowptr[(size_t)b * (p.T + 1) + i + 1];
        const int* __restrict__ col = p.col + (size_t)b * p.cap;
        const unsigned char* __restrict__ val = p.val + (size_t)b * p.cap;
        const float dn = p.denom[n];
        for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
            float acc = p.self_loop ? p.S[(size_t)n * p.H + h] : 0.f;
            for (int e = e0; e < e1; ++e) {
                const int j = col[e], v = val[e];
                if (is_fwd(v)) {
                    if (edge_kept(p.keep_f, p.rng, p.edge_keep, p.layer, 0u, b, i, j, p.T))
                        acc += p.F[((size_t)b * p.T + j) * p.H + h];
                } else if (is_rev(v) && !p.directed) {
                    if (edge_kept(p.keep_r, p.rng, p.edge_keep, p.layer, 1u, b, i, j, p.T))
                        acc += p.R[((size_t)b * p.T + j) * p.H + h];
                }
            }
            float o = fmaxf(acc / dn, 0.f);
            const size_t idx = (size_t)n * p.H + h;
            if (p.drop_mask != nullptr) {
                o *= p.drop_mask[idx];
            } else if (philox_drop) {
                const unsigned long long seed = p.rng[0], step = p.rng[1];
                const Philox4 r = philox4x32((uint32_t)idx, (uint32_t)(idx >> 32), 0xD7090000u | p.layer, (uint32_t)step,
                                             (uint32_t)seed, (uint32_t)(seed >> 32));
                o = (u01(r.x) >= p.drop_p) ? o * scale : 0.f;
            }
            p.out[idx] = o;
        }
    }
}

// d(out)/d(z) recovered from the forward's output: out != 0 <=> the element was kept AND z > 0
__device__ __forceinline__ float upstream(const Agg3Params& p, size_t idx, float scale, float inv_dn) {
    const float o = p.outp[idx];
    if (o == 0.f) return 0.f;
    const float m = p.drop_mask != nullptr ? p.drop_mask[idx] : scale;
    return p.gout[idx] * m * inv_dn;
}

__global__ void __launch_bounds__(kThreads) agg3_bwd_kernel(const Agg3Params p) {
    GPT_PDL_ENTER();
    const int N = p.B * p.T;
    const bool philox_drop = p.drop_mask == nullptr && p.drop_p > 0.f && p.rng != nullptr;
    const float scale = philox_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        const int b = n / p.T, j = n - b * p.T;
        if (!observable(p.flags[n])) {
            for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
                const size_t idx = (size_t)n * p.H + h;
                p.dF[idx] = 0.f;
                p.dR[idx] = 0.f;
                p.dS[idx] = 0.f;
            }
            continue;
        }
        const int e0 = p.rowptr[(size_t)b * (p.T + 1) + j], e1 = p.rowptr[(size_t)b * (p.T + 1) + j + 1];
        const int* __restrict__ col = p.col + (size_t)b * p.cap;
        const unsigned char* __restrict__ val = p.val + (size_t)b * p.cap;
        const float inv_self = 1.f / p.denom[n];
        for (int h = threadIdx.x; h < p.H; h += blockDim.x) {
            const size_t idx = (size_t)n * p.H + h;
            float df = 0.f, dr = 0.f;
            for (int e = e0; e < e1; ++e) {
                const int i = col[e], v = val[e];
                const size_t nb = ((size_t)b * p.T + i) * p.H + h;
                if (is_rev(v)) {
                    // i is j's parent: its row gathered F_j through entry [i, j] of the parent->child matrix
                    if (edge_kept(p.keep_f, p.rng, p.edge_keep, p.layer, 0u, b, i, j, p.T))
                        df += upstream(p, nb, scale, 1.f / p.denom[(size_t)b * p.T + i]);
                } else if (is_fwd(v) && !p.directed) {
                    // i is a child of j: its row gathered R_j through entry [i, j] of the child->parent matrix
                    if (edge_kept(p.keep_r, p.rng, p.edge_keep, p.layer, 1u, b, i, j, p.T))
                        dr += upstream(p, nb, scale, 1.f / p.denom[(size_t)b * p.T + i]);
                }
            }
            p.dF[idx] = df;
            p.dR[idx] = dr;
            p.dS[idx] = p.self_loop ? upstream(p, idx, scale, inv_self) : 0.f;
        }
    }
}

// ---- the same aggregation, one WARP per token row and 128-bit columns (H % 4 == 0) ------------------------------------
// A row moves a few KB behind a chain of dependent loads (flags -> rowptr -> col / val -> the neighbours' rows): what
// matters is how many rows are in flight, so a 128-thread CTA serves four rows at once and a lane holds four columns.
constexpr int kAggWarps = kThreads / 32;

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__global__ void __launch_bounds__(kThreads) agg3_fwd4_kernel(const Agg3Params p) {
    GPT_PDL_ENTER();
    const int N = p.B * p.T, hq = p.H >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool philox_drop = p.drop_mask == nullptr && p.drop_p > 0.f && p.rng != nullptr;
    const float scale = philox_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = blockIdx.x * kAggWarps + warp; n < N; n += gridDim.x * kAggWarps) {
        float4* __restrict__ out4 = reinterpret_cast<float4*>(p.out + (size_t)n * p.H);
        if (!observable(p.flags[n])) {                 // warp-uniform
            for (int q = lane; q < hq; q += 32) out4[q] = zero;
            continue;
        }
        const int b = n / p.T, i = n - b * p.T;
        const int e0 = p.rowptr[(size_t)b * (p.T + 1) + i], e1 = p.rowptr[(size_t)b * (p.T + 1) + i + 1];
        const int* __restrict__ col = p.col + (size_t)b * p.cap;
        const unsigned char* __restrict__ val = p.val + (size_t)b * p.cap;
        const float dn = p.denom[n];
        for (int q = lane; q < hq; q += 32) {
            float4 acc = p.self_loop ? reinterpret_cast<const float4*>(p.S + (size_t)n * p.H)[q] : zero;
            for (int e = e0; e < e1; ++e) {
                const int j = col[e], v = val[e];
                if (is_fwd(v)) {
                    if (edge_kept(p.keep_f, p.rng, p.edge_keep, p.layer, 0u, b, i, j, p.T))
                        acc = f4_add(acc, reinterpret_cast<const float4*>(p.F + ((size_t)b * p.T + j) * p.H)[q]);
                } else if (is_rev(v) && !p.directed) {
                    if (edge_kept(p.keep_r, p.rng, p.edge_keep, p.layer, 1u, b, i, j, p.T))
                        acc = f4_add(acc, reinterpret_cast<const float4*>(p.R + ((size_t)b * p.T + j) * p.H)[q]);
                }
            }
            float o[4] = {fmaxf(acc.x / dn, 0.f), fmaxf(acc.y / dn, 0.f), fmaxf(acc.z / dn, 0.f), fmaxf(acc.w / dn, 0.f)};
            const size_t idx0 = (size_t)n * p.H + 4 * q;
            if (p.drop_mask != nullptr) {
                const float4 m = reinterpret_cast<const float4*>(p.drop_mask + idx0)[0];
                o[0] *= m.x; o[1] *= m.y; o[2] *= m.z; o[3] *= m.w;
            } else if (philox_drop) {
                const unsigned long long seed = p.rng[0], step = p.rng[1];
#pragma unroll
                for (int c = 0; c < 4; ++c) {           // the element's own stream, as the scalar kernel draws it
                    const size_t idx = idx0 + c;
                    const Philox4 r = philox4x32((uint32_t)idx, (uint32_t)(idx >> 32), 0xD7090000u | p.layer, (uint32_t)step,
                                                 (uint32_t)seed, (uint32_t)(seed >> 32));
                    o[c] = (u01(r.x) >= p.drop_p) ? o[c] * scale : 0.f;
                }
            }
            out4[q] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

__device__ __forceinline__ float4 upstream4(const Agg3Params& p, size_t idx0, float scale, float inv_dn) {
    const float4 o = reinterpret_cast<const float4*>(p.outp + idx0)[0];
    const float4 g = reinterpret_cast<const float4*>(p.gout + idx0)[0];
    float4 m = make_float4(scale, scale, scale, scale);
    if (p.drop_mask != nullptr) m = reinterpret_cast<const float4*>(p.drop_mask + idx0)[0];
    return make_float4(o.x == 0.f ? 0.f : g.x * m.x * inv_dn, o.y == 0.f ? 0.f : g.y * m.y * inv_dn,
                       o.z == 0.f ? 0.f : g.z * m.z * inv_dn, o.w == 0.f ? 0.f : g.w * m.w * inv_dn);
}

__global__ void __launch_bounds__(kThreads) agg3_bwd4_kernel(const Agg3Params p) {
    GPT_PDL_ENTER();
    const int N = p.B * p.T, hq = p.H >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool philox_drop = p.drop_mask == nullptr && p.drop_p > 0.f && p.rng != nullptr;
    const float scale = philox_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = blockIdx.x * kAggWarps + warp; n < N; n += gridDim.x * kAggWarps) {
        float4* __restrict__ dF4 = reinterpret_cast<float4*>(p.dF + (size_t)n * p.H);
        float4* __restrict__ dR4 = reinterpret_cast<float4*>(p.dR + (size_t)n * p.H);
        float4* __restrict__ dS4 = reinterpret_cast<float4*>(p.dS + (size_t)n * p.H);
        if (!observable(p.flags[n])) {                 // warp-uniform
            for (int q = lane; q < hq; q += 32) dF4[q] = dR4[q] = dS4[q] = zero;
            continue;
        }
        const int b = n / p.T, j = n - b * p.T;
        const int e0 = p.rowptr[(size_t)b * (p.T + 1) + j], e1 = p.rowptr[(size_t)b * (p.T + 1) + j + 1];
        const int* __restrict__ col = p.col + (size_t)b * p.cap;
        const unsigned char* __restrict__ val = p.val + (size_t)b * p.cap;
        const float inv_self = 1.f / p.denom[n];
        for (int q = lane; q < hq; q += 32) {
            float4 df = zero, dr = zero;
            for (int e = e0; e < e1; ++e) {
                const int i = col[e], v = val[e];
                const size_t nb = ((size_t)b * p.T + i) * p.H + 4 * q;
                if (is_rev(v)) {            // i is j's parent: its row gathered F_j through entry [i, j] of the parent->child matrix
                    if (edge_kept(p.keep_f, p.rng, p.edge_keep, p.layer, 0u, b, i, j, p.T))
                        df = f4_add(df, upstream4(p, nb, scale, 1.f / p.denom[(size_t)b * p.T + i]));
                } else if (is_fwd(v) && !p.directed) {   // i is a child of j: its row gathered R_j through entry [i, j]
                    if (edge_kept(p.keep_r, p.rng, p.edge_keep, p.layer, 1u, b, i, j, p.T))
                        dr = f4_add(dr, upstream4(p, nb, scale, 1.f / p.denom[(size_t)b * p.T + i]));
                }
            }
            dF4[q] = df;
            dR4[q] = dr;
            dS4[q] = p.self_loop ? upstream4(p, (size_t)n * p.H + 4 * q, scale, inv_self) : zero;
        }
    }
}

static bool agg3_vec_ok(const Agg3Params& p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.F) | reinterpret_cast<uintptr_t>(p.R) | reinterpret_cast<uintptr_t>(p.S) |
                        reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.gout) |
                        reinterpret_cast<uintptr_t>(p.outp) | reinterpret_cast<uintptr_t>(p.dF) |
                        reinterpret_cast<uintptr_t>(p.dR) | reinterpret_cast<uintptr_t>(p.dS) |
                        reinterpret_cast<uintptr_t>(p.drop_mask);
    return p.H % 4 == 0 && (a & 15) == 0;
}

// dense [B,T,T] copy of the Philox edge-keep decisions of one layer and direction (tests feed it to the oracle)
__global__ void edge_keep_dense_kernel(const unsigned long long* __restrict__ rng, int B, int T, unsigned layer, unsigned dir,
                                       float keep_prob, unsigned char* __restrict__ out) {
    GPT_PDL_ENTER();
    const size_t total = (size_t)B * T * T;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(k % T);
        const size_t bi = k / T;
        const int i = (int)(bi % T), b = (int)(bi / T);
        out[k] = edge_kept(nullptr, rng, keep_prob, layer, dir, b, i, j, T) ? 1 : 0;
    }
}

// per-token Bernoulli(keep_prop) for relation forgetting, one draw per direction (gcn.py:451-470)
__global__ void token_keep_kernel(const unsigned long long* __restrict__ rng, int N, unsigned layer, float keep_prop,
                                  unsigned char* __restrict__ keep_f, unsigned char* __restrict__ keep_r) {
    GPT_PDL_ENTER();
    const unsigned long long seed = rng[0], step = rng[1];
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const Philox4 r = philox4x32((uint32_t)n, 0u, 0xF0460000u | layer, (uint32_t)step, (uint32_t)seed,
                                     (uint32_t)(seed >> 32));
        keep_f[n] = u01(r.x) < keep_prop ? 1 : 0;
        keep_r[n] = u01(r.y) < keep_prop ? 1 : 0;
    }
}

// out[c] += sum_r a[r, c]   (bias gradient of the shared projection: column sums of dZ)
__global__ void __launch_bounds__(256) colsum_acc_kernel(const float* __restrict__ a, long long rows, int cols, int chunk,
                                                         float* __restrict__ out, const int* __restrict__ count) {
    GPT_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    if (count != nullptr && *count < rows) rows = *count;       // only the compacted observable rows (gpt_live_rows)
    const long long r0 = (long long)blockIdx.y * chunk;
    const long long r1 = r0 + chunk < rows ? r0 + chunk : rows;
    if (r0 >= r1) return;
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += a[(size_t)r * cols + c];
    atomicAdd(out + c, s);
}

inline unsigned row_grid(long long rows, int threads = kThreads) {
    long long cap = 148LL * (2048 / threads);   // resident CTAs per SM x SM count: one wave, grid-stride beyond
    if (const char* e = getenv("GPT_K10_MAX_CTAS")) {      // tuning / test knob: CTAs per launch
        const long long v = atoll(e);
        if (v > 0) cap = v;
    }
    return (unsigned)(rows < 1 ? 1 : (rows < cap ? rows : cap));
}

}  // namespace

constexpr size_t kPipeSmemMax = 110 * 1024;     // two CTAs per SM
// pipelined row walkers: two resident CTAs per SM, so that a CTA sees several rows and its prefetch warp has work to hide
static unsigned pipe_grid(long long rows) {
    long long cap = 148LL * 2;
    if (const char* e = getenv("GPT_K10_MAX_CTAS")) {
        const long long v = atoll(e);
        if (v > 0) cap = v;
    }
    return (unsigned)(rows < 1 ? 1 : (rows < cap ? rows : cap));
}

static bool mix_vec_ok(const MixParams& p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.Z) | reinterpret_cast<uintptr_t>(p.bias) | reinterpret_cast<uintptr_t>(p.F) |
                        reinterpret_cast<uintptr_t>(p.R) | reinterpret_cast<uintptr_t>(p.S) | reinterpret_cast<uintptr_t>(p.dZ);
    return p.H % 4 == 0 && (a & 15) == 0;
}

// perm / count: NULL = every token row; else the compact row list of gpt_live_rows (Z is then [count, D*H], compact)
extern "C" int gpt_relmix_fwd_rows(const float* Z, const float* bias, const float* E, const int64_t* deprel,
                                   const uint8_t* flags, const uint8_t* keep_f, const uint8_t* keep_r, const int32_t* perm,
                                   const int32_t* count, int N, int D, int H, int deep, float* F, float* R, float* S,
                                   void* stream) {
    GPT_CHECK_ARG(Z && E && deprel && flags && F && R && S && N >= 0 && D >= 1 && H >= 1);    // bias NULL: already in Z
    GPT_CHECK_ARG((perm == nullptr) == (count == nullptr));
    if (N == 0) return GPT_OK;
    MixParams p{};
    p.Z = Z; p.bias = bias; p.E = E; p.deprel = reinterpret_cast<const long long*>(deprel); p.flags = flags;
    p.keep_f = keep_f; p.keep_r = keep_r; p.F = F; p.R = R; p.S = S; p.N = N; p.D = D; p.H = H; p.deep = deep;
    p.perm = perm; p.count = count;
    const size_t smem4 = ((size_t)((3 * D + 3) & ~3) + (size_t)(kMixThreads / 32) * 3 * H) * sizeof(float);
    const size_t smem_pipe = smem4 + ((size_t)((3 * D + 3) & ~3) + 2 * (size_t)D * H) * sizeof(float);
    if (perm != nullptr && mix_vec_ok(p) && smem_pipe <= kPipeSmemMax) {
        if (int a = gpt_smem_opt_in(relmix_fwd_pipe_kernel, smem_pipe)) return a;
        gpt_launch(relmix_fwd_pipe_kernel, dim3(pipe_grid(N)), dim3(kPipeThreads), smem_pipe, (cudaStream_t)stream, p);
        return gpt_launch_status();
    }
    if (mix_vec_ok(p) && smem4 <= 48 * 1024) {
        gpt_launch(relmix_fwd4_kernel, dim3(row_grid(N, kMixThreads)), dim3(kMixThreads), smem4, (cudaStream_t)stream, p);
        return gpt_launch_status();
    }
    if ((size_t)3 * D * sizeof(float) > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    gpt_launch(relmix_fwd_kernel, dim3(row_grid(N)), dim3(kThreads), (size_t)3 * D * sizeof(float), (cudaStream_t)stream, p);
    return gpt_launch_status();
}

extern "C" int gpt_relmix_fwd(const float* Z, const float* bias, const float* E, const int64_t* deprel, const uint8_t* flags,
                              const uint8_t* keep_f, const uint8_t* keep_r, int N, int D, int H, int deep, float* F,
                              float* R, float* S, void* stream) {
    return gpt_relmix_fwd_rows(Z, bias, E, deprel, flags, keep_f, keep_r, nullptr, nullptr, N, D, H, deep, F, R, S, stream);
}

extern "C" int gpt_relmix_bwd_rows(const float* Z, const float* bias, const float* E, const int64_t* deprel,
                                   const uint8_t* flags, const uint8_t* keep_f, const uint8_t* keep_r, const int32_t* perm,
                                   const int32_t* count, const float* dF, const float* dR, const float* dS, int N, int D,
                                   int H, int deep, float* dZ, float* dE, void* stream) {
    GPT_CHECK_ARG(Z && E && deprel && flags && dF && dR && dS && dZ && dE && N >= 0 && D >= 1 && H >= 1);
    GPT_CHECK_ARG((perm == nullptr) == (count == nullptr));
    if (N == 0) return GPT_OK;
    MixParams p{};
    p.Z = Z; p.bias = bias; p.E = E; p.deprel = reinterpret_cast<const long long*>(deprel); p.flags = flags;
    p.keep_f = keep_f; p.keep_r = keep_r; p.F = const_cast<float*>(dF); p.R = const_cast<float*>(dR);
    p.S = const_cast<float*>(dS); p.dZ = dZ; p.dE = dE; p.N = N; p.D = D; p.H = H; p.deep = deep;
    p.perm = perm; p.count = count;
    const size_t smem4 = ((size_t)((4 * D + 3) & ~3) + (size_t)3 * H) * sizeof(float);
    const size_t smem_pipe =
        (2 * (size_t)D * H + (size_t)((D + 3) & ~3) + 2 * ((size_t)((3 * D + 3) & ~3) + 3 * H)) * sizeof(float);
    if (perm != nullptr && mix_vec_ok(p) && smem_pipe <= kPipeSmemMax) {
        if (int a = gpt_smem_opt_in(relmix_bwd_pipe_kernel, smem_pipe)) return a;
        gpt_launch(relmix_bwd_pipe_kernel, dim3(pipe_grid(N)), dim3(kPipeThreads), smem_pipe, (cudaStream_t)stream, p);
        return gpt_launch_status();
    }
    if (mix_vec_ok(p) && smem4 <= 48 * 1024) {
        gpt_launch(relmix_bwd4_kernel, dim3(row_grid(N, kMixThreads)), dim3(kMixThreads), smem4, (cudaStream_t)stream, p);
        return gpt_launch_status();
    }
    const size_t smem = ((size_t)4 * D + 3 * H) * sizeof(float);
    if (smem > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    gpt_launch(relmix_bwd_kernel, dim3(row_grid(N)), dim3(kThreads), smem, (cudaStream_t)stream, p);
    return gpt_launch_status();
}

extern "C" int gpt_relmix_bwd(const float* Z, const float* bias, const float* E, const int64_t* deprel, const uint8_t* flags,
                              const uint8_t* keep_f, const uint8_t* keep_r, const float* dF, const float* dR,
                              const float* dS, int N, int D, int H, int deep, float* dZ, float* dE, void* stream) {
    return gpt_relmix_bwd_rows(Z, bias, E, deprel, flags, keep_f, keep_r, nullptr, nullptr, dF, dR, dS, N, D, H, deep, dZ, dE,
                               stream);
}

// flags [N] -> perm [N], inv [N], live [N], count [1] (see live_rows_kernel)
extern "C" int gpt_live_rows(const uint8_t* flags, int N, int32_t* perm, int32_t* inv, uint8_t* live, int32_t* count,
                             void* stream) {
    GPT_CHECK_ARG(flags && perm && inv && live && count && N >= 0);
    gpt_launch(live_rows_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, flags, N, perm, inv, live, count);
    return gpt_launch_status();
}

extern "C" int gpt_gather_rows(const float* x, const int32_t* perm, const int32_t* count, int N, int K, float* out,
                               void* stream) {
    GPT_CHECK_ARG(x && perm && count && out && N >= 0 && K >= 1);
    if (N == 0) return GPT_OK;
    gpt_launch(gather_rows_kernel, dim3(row_grid(N, 256)), dim3(256), 0, (cudaStream_t)stream, x, perm, count, N, K, out);
    return gpt_launch_status();
}

extern "C" int gpt_scatter_rows(const float* dxc, const int32_t* inv, int N, int K, float* dx, void* stream) {
    GPT_CHECK_ARG(dxc && inv && dx && N >= 0 && K >= 1);
    if (N == 0) return GPT_OK;
    gpt_launch(scatter_rows_kernel, dim3(row_grid(N, 256)), dim3(256), 0, (cudaStream_t)stream, dxc, inv, N, K, dx);
    return gpt_launch_status();
}

extern "C" int gpt_diagmix_fwd(const float* x, const float* E, const int64_t* deprel, const uint8_t* flags, int N, int H,
                               float* F, float* R, float* S, void* stream) {
    GPT_CHECK_ARG(x && E && deprel && flags && F && R && S && N >= 0 && H >= 1);
    if (N == 0) return GPT_OK;
    DiagParams p{};
    p.x = x; p.E = E; p.deprel = reinterpret_cast<const long long*>(deprel); p.flags = flags; p.F = F; p.R = R; p.S = S;
    p.N = N; p.H = H;
    gpt_launch(diagmix_fwd_kernel, dim3(row_grid(N)), dim3(kThreads), 0, (cudaStream_t)stream, p);
    return gpt_launch_status();
}

extern "C" int gpt_diagmix_bwd(const float* x, const float* E, const int64_t* deprel, const uint8_t* flags, const float* dF,
                               const float* dR, const float* dS, int N, int H, float* dx, float* dE, void* stream) {
    GPT_CHECK_ARG(x && E && deprel && flags && dF && dR && dS && dx && dE && N >= 0 && H >= 1);
    if (N == 0) return GPT_OK;
    if ((size_t)H * sizeof(float) > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    DiagParams p{};
    p.x = x; p.E = E; p.deprel = reinterpret_cast<const long long*>(deprel); p.flags = flags;
    p.F = const_cast<float*>(dF); p.R = const_cast<float*>(dR); p.S = const_cast<float*>(dS); p.dx = dx; p.dE = dE;
    p.N = N; p.H = H;
    gpt_launch(diagmix_bwd_kernel, dim3(row_grid(N)), dim3(kThreads), (size_t)H * sizeof(float), (cudaStream_t)stream, p);
    return gpt_launch_status();
}

static int fill_agg3(Agg3Params& p, const int32_t* rowptr, const int32_t* col, const uint8_t* val, const float* denom,
                     const uint8_t* flags, const uint8_t* keep_f, const uint8_t* keep_r, float edge_keep, const void* rng,
                     unsigned layer, int directed, int self_loop, float drop_p, const float* drop_mask, int B, int T, int H) {
    GPT_CHECK_ARG(rowptr && col && val && denom && flags && B >= 0 && T >= 1 && H >= 1);
    GPT_CHECK_ARG(edge_keep >= 0.f && drop_p >= 0.f && drop_p < 1.f && layer < 0x8000u);
    GPT_CHECK_ARG(!(edge_keep < 1.f && !(keep_f && keep_r) && rng == nullptr));
    p.rowptr = rowptr; p.col = col; p.val = val; p.denom = denom; p.flags = flags; p.keep_f = keep_f; p.keep_r = keep_r;
    p.rng = static_cast<const unsigned long long*>(rng); p.drop_mask = drop_mask; p.B = B; p.T = T; p.H = H; p.cap = 3 * T;
    p.directed = directed; p.self_loop = self_loop; p.layer = layer; p.edge_keep = edge_keep; p.drop_p = drop_p;
    return GPT_OK;
}

extern "C" int gpt_agg3_fwd(const float* F, const float* R, const float* S, const int32_t* rowptr, const int32_t* col,
                            const uint8_t* val, const float* denom, const uint8_t* flags, const uint8_t* keep_f,
                            const uint8_t* keep_r, float edge_keep, const void* rng_state, unsigned layer, int directed,
                            int self_loop, float drop_p, const float* drop_mask, int B, int T, int H, float* out,
                            void* stream) {
    GPT_CHECK_ARG(F && R && S && out);
    Agg3Params p{};
    int rc = fill_agg3(p, rowptr, col, val, denom, flags, keep_f, keep_r, edge_keep, rng_state, layer, directed, self_loop,
                       drop_p, drop_mask, B, T, H);
    if (rc != GPT_OK) return rc;
    if (B == 0) return GPT_OK;
    p.F = F; p.R = R; p.S = S; p.out = out;
    if (agg3_vec_ok(p)) {
        const long long rows = (long long)B * T;
        gpt_launch(agg3_fwd4_kernel, dim3(row_grid((rows + kAggWarps - 1) / kAggWarps)), dim3(kThreads), 0, (cudaStream_t)stream, p);
        return gpt_launch_status();
    }
    gpt_launch(agg3_fwd_kernel, dim3(row_grid((long long)B * T)), dim3(kThreads), 0, (cudaStream_t)stream, p);
    return gpt_launch_status();
}

extern "C" int gpt_agg3_bwd(const float* gout, const float* out, const int32_t* rowptr, const int32_t* col,
                            const uint8_t* val, const float* denom, const uint8_t* flags, const uint8_t* keep_f,
                            const uint8_t* keep_r, float edge_keep, const void* rng_state, unsigned layer, int directed,
                            int self_loop, float drop_p, const float* drop_mask, int B, int T, int H, float* dF, float* dR,
                            float* dS, void* stream) {
    GPT_CHECK_ARG(gout && out && dF && dR && dS);
    Agg3Params p{};
    int rc = fill_agg3(p, rowptr, col, val, denom, flags, keep_f, keep_r, edge_keep, rng_state, layer, directed, self_loop,
                       drop_p, drop_mask, B, T, H);
    if (rc != GPT_OK) return rc;
    if (B == 0) return GPT_OK;
    p.gout = gout; p.outp = out; p.dF = dF; p.dR = dR; p.dS = dS;
    if (agg3_vec_ok(p)) {
        const long long rows = (long long)B * T;
        gpt_launch(agg3_bwd4_kernel, dim3(row_grid((rows + kAggWarps - 1) / kAggWarps)), dim3(kThreads), 0, (cudaStream_t)stream, p);
        return gpt_launch_status();
    }
    gpt_launch(agg3_bwd_kernel, dim3(row_grid((long long)B * T)), dim3(kThreads), 0, (cudaStream_t)stream, p);
    return gpt_launch_status();
}

extern "C" int gpt_edge_keep_dense(const void* rng_state, int B, int T, unsigned layer, int dir, float keep_prob,
                                   uint8_t* out, void* stream) {
    GPT_CHECK_ARG(rng_state && out && B >= 0 && T >= 1 && (dir == 0 || dir == 1) && layer < 0x8000u);
    if (B == 0) return GPT_OK;
    const size_t total = (size_t)B * T * T;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gpt_launch(edge_keep_dense_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
               static_cast<const unsigned long long*>(rng_state), B, T, layer, (unsigned)dir, keep_prob, out);
    return gpt_launch_status();
}

extern "C" int gpt_relation_keep_tokens(const void* rng_state, int N, unsigned layer, float keep_prop, uint8_t* keep_f,
                                        uint8_t* keep_r, void* stream) {
    GPT_CHECK_ARG(rng_state && keep_f && keep_r && N >= 0 && layer < 0x8000u);
    if (N == 0) return GPT_OK;
    int blocks = (N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gpt_launch(token_keep_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
               static_cast<const unsigned long long*>(rng_state), N, layer, keep_prop, keep_f, keep_r);
    return gpt_launch_status();
}

// out[c] += sum over rows r < min(rows, *count) of a[r, c]   (count == NULL: every row)
extern "C" int gpt_colsum_acc_rows(const float* a, long long rows, int cols, const int32_t* count, float* out, void* stream) {
    GPT_CHECK_ARG(a && out && rows >= 0 && cols >= 1);
    if (rows == 0) return GPT_OK;
    const int chunk = 64;
    const long long row_blocks = (rows + chunk - 1) / chunk;
    GPT_CHECK_ARG(row_blocks <= 65535);
    gpt_launch(colsum_acc_kernel, dim3((unsigned)((cols + 255) / 256), (unsigned)row_blocks), dim3(256), 0,
               (cudaStream_t)stream, a, rows, cols, chunk, out, count);
    return gpt_launch_status();
}

extern "C" int gpt_colsum_acc(const float* a, long long rows, int cols, float* out, void* stream) {
    return gpt_colsum_acc_rows(a, rows, cols, nullptr, out, stream);
}
