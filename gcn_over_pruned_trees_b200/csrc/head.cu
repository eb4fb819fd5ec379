// K6 -- classifier head: out_mlp + classifier + loss, forward and backward, in two launches.
//
// Replaces, for one batch of pooled sentence vectors [B, 3H] (the output of K4):
//   out_mlp    Linear(3H,H)+ReLU, (mlp_layers-1) x Linear(H,H)+ReLU       /root/reference/model/gcn.py:64-68,122
//   classifier Linear(H, num_class)                                        model/gcn.py:21,29
//   loss       CrossEntropy(mean) + pooling_l2 * mean_b sum_h h_out^2      model/trainer.py:94-100
// and the autograd of all of it (loss.backward(), train.py:221).  In the reference this is ~45 ATen launches per
// step on [50, <=600] operands -- pure launch latency.  Here:
//   head_fwd_bwd_kernel : one CTA per sentence (R sentences when B is large).  Every layer is a matrix-vector
//       product against weights that all CTAs stream from L2 at the same time: warp-per-output-row dot products with
//       128-bit loads for the forward, thread-per-column sums for the data gradient.  Activations never leave shared
//       memory; the kernel emits logits, the per-sentence loss, d(loss)/d(pooled) and what the weight gradient needs.
//   head_wgrad_kernel   : dW_l = dpre_l^T . in_l and db_l for every layer as 32x64 tiles of one flat grid, each tile
//       summing over the whole batch (no atomics: deterministic), plus one CTA that adds the per-sentence losses in
//       a fixed order.
#include "gpt_common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

constexpr int kHeadThreads = 512;
constexpr int kMaxMlp = 4;
constexpr int kDgradPartCols = 2048;  // floats of split-N partial sums per sentence (G * K <= 4 * kHeadThreads)

struct HeadParams {
    const float* pooled;        // [B, 3H]
    const long long* labels;    // [B]
    const float* w[kMaxMlp];    // w[0] [H, 3H], w[l>0] [H, H]   (nn.Linear layout)
    const float* b[kMaxMlp];    // [H]
    const float* wc;            // [C, H]
    const float* bc;            // [C]
    int B, H, C, n_mlp, train, prefetch;
    float pooling_l2, inv_B;
    float* logits;              // [B, C]
    float* loss_rows;           // [B]  CE_b / B + pooling_l2 * |h_out_b|^2 / B
    float* acts;                // [B, n_mlp, H]  post-ReLU activations          (train)
    float* dacts;               // [B, n_mlp, H]  d loss / d pre-activation      (train)
    float* dlogits;             // [B, C]                                        (train)
    float* dpooled;             // [B, 3H]                                       (train)
};

// volatile: ptxas keeps these in program order, i.e. all loads of a batch are issued before the first consumer
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// s_out[r][n] = act(sum_k s_in[r][k] * W[n][k] + bias[n]);  K % 4 == 0; one warp per 4 output rows of W.
// The CTAs of a cluster share one sentence group: they take the 4-row groups round-robin (each streams 1/CS of W) and
// store every result into the shared memory of ALL ranks (DSMEM), so after the cluster barrier each holds the layer.
template <int R, bool RELU>
__device__ __forceinline__ void dense_fwd(cg::cluster_group& cluster, const float* __restrict__ W,
                                          const float* __restrict__ bias, const float* s_in, int ld_in, float* s_out,
                                          int ld_out, int N, int K) {
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int cs = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int gwarp = (int)(threadIdx.x >> 5) * cs + crank;          // interleave ranks: balanced for any N
    const int K4 = K >> 2;
    for (int n0 = gwarp * 4; n0 < N; n0 += nw * cs * 4) {
        float acc[4][R];
        const float4* wr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            wr[j] = reinterpret_cast<const float4*>(W + (size_t)min(n0 + j, N - 1) * K);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[j][r] = 0.f;
        }
        // U x 4 independent 128-bit loads are issued before the first use (addresses clamped, tails zeroed through
        // the activation operand): the loop is bound by round trips to L2, so loads in flight are what matters
        constexpr int U = 3;
        for (int k4 = lane; k4 < K4; k4 += 32 * U) {
            float4 wv[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = min(k4 + 32 * u, K4 - 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) wv[u][j] = ld_nc_f4(wr[j] + kk);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool in = k4 + 32 * u < K4;
                const int kk = min(k4 + 32 * u, K4 - 1);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float4 x = reinterpret_cast<const float4*>(s_in + r * ld_in)[kk];
                    if (!in) x = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[j][r] += wv[u][j].x * x.x + wv[u][j].y * x.y + wv[u][j].z * x.z + wv[u][j].w * x.w;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float v = warp_sum_f(acc[j][r]);
                if (lane == j * R + r && n0 + j < N) {
                    float o = v + bias[n0 + j];
                    o = RELU ? fmaxf(o, 0.f) : o;
                    for (int q = 0; q < cs; ++q) cluster.map_shared_rank(s_out, q)[r * ld_out + n0 + j] = o;
                }
            }
        }
    }
}

// s_din[r][k] = sum_n s_dout[r][n] * W[n][k]; a thread owns 4 consecutive k (one 128-bit column of W) and a residue
// class of n; the G classes are added through s_part.  Rows of W whose dout is zero for every sentence of the CTA
// (ReLU-dead units: about half) are skipped.
// In a cluster every rank owns a contiguous range of the 128-bit columns (1/CS of W) and, unless `broadcast` is off
// (last layer: the result goes straight to global memory), stores its part of s_din into every rank.
template <int R>
__device__ __forceinline__ void dense_dgrad(cg::cluster_group& cluster, const float* __restrict__ W,
                                            const float* s_dout, int ld_dout, float* s_part, float* s_din,
                                            int ld_din, int N, int K, bool broadcast, int* k_lo_out, int* k_hi_out) {
    const int K4 = K >> 2;
    const int cs = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int k4_lo = (int)(((long long)K4 * crank) / cs), k4_hi = (int)(((long long)K4 * (crank + 1)) / cs);
    const int K4c = k4_hi - k4_lo;                          // this rank's columns
    *k_lo_out = k4_lo * 4;
    *k_hi_out = k4_hi * 4;
    const int G = (K4c > 0 && K4c <= (int)blockDim.x) ? min((int)blockDim.x / K4c, kDgradPartCols / (4 * K4c)) : 1;
    const int Gc = G < 1 ? 1 : G;
    for (int item = threadIdx.x; item < Gc * K4c; item += blockDim.x) {
        const int g = item / K4c, k4 = k4_lo + (item - g * K4c);
        float4 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* wc = reinterpret_cast<const float4*>(W) + k4;
        constexpr int U = 8;                                    // rows of W in flight per thread
        for (int n = g; n < N; n += Gc * U) {
            float4 wv[U];
            float d[U][R];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int nn = n + u * Gc;
#pragma unroll
                for (int r = 0; r < R; ++r) d[u][r] = nn < N ? s_dout[r * ld_dout + nn] : 0.f;
                wv[u] = ld_nc_f4(wc + (size_t)min(nn, N - 1) * K4);     // unconditional: keeps the batch in flight
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r].x += d[u][r] * wv[u].x; acc[r].y += d[u][r] * wv[u].y;
                    acc[r].z += d[u][r] * wv[u].z; acc[r].w += d[u][r] * wv[u].w;
                }
        }
        if (Gc == 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (broadcast) {
                    for (int q = 0; q < cs; ++q)
                        reinterpret_cast<float4*>(cluster.map_shared_rank(s_din, q) + r * ld_din)[k4] = acc[r];
                } else {
                    reinterpret_cast<float4*>(s_din + r * ld_din)[k4] = acc[r];
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                reinterpret_cast<float4*>(s_part + (size_t)(g * R + r) * (4 * K4c))[k4 - k4_lo] = acc[r];
        }
    }
    __syncthreads();
    if (Gc > 1) {
        const int Kc = 4 * K4c;
        for (int i = threadIdx.x; i < R * Kc; i += blockDim.x) {
            const int r = i / Kc, kk = i - r * Kc;
            float s = 0.f;
            for (int g = 0; g < Gc; ++g) s += s_part[(size_t)(g * R + r) * Kc + kk];
            const int k = k4_lo * 4 + kk;
            if (broadcast) {
                for (int q = 0; q < cs; ++q) cluster.map_shared_rank(s_din, q)[r * ld_din + k] = s;
            } else {
                s_din[r * ld_din + k] = s;
            }
        }
    }
    cluster.sync();
}

template <int R>
__global__ void __launch_bounds__(kHeadThreads, 1)
head_fwd_bwd_kernel(const HeadParams p) {
    GPT_PDL_ENTER();
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, C = p.C, K0 = 3 * p.H, L = p.n_mlp;
    const int Cp = (C + 3) & ~3;
    float* s_in = smem;                          // [R][K0]      pooled rows, later d(pooled)
    float* s_h = s_in + R * K0;                  // [L][R][H]    post-ReLU activations
    float* s_logit = s_h + L * R * H;            // [R][Cp]
    float* s_dl = s_logit + R * Cp;              // [R][Cp]
    float* s_d0 = s_dl + R * Cp;                 // [R][H]
    float* s_d1 = s_d0 + R * H;                  // [R][H]
    float* s_part = s_d1 + R * H;                // [R][max(kDgradPartCols, K0)]
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int b0 = (int)(blockIdx.x / cs) * R;
    const int tid = threadIdx.x;

    // warm L2: together the CTAs touch every weight line once, so the DRAM latency is paid once and in parallel
    // instead of once per pass of every layer
    if (p.prefetch) {
        const size_t gtid = (size_t)blockIdx.x * blockDim.x + tid, gsize = (size_t)gridDim.x * blockDim.x;
        for (int l = 0; l <= L; ++l) {
            const float* base = l < L ? p.w[l] : p.wc;
            const size_t lines = ((size_t)(l < L ? H : C) * (l == 0 ? K0 : H) * sizeof(float) + 127) / 128;
            for (size_t i = gtid; i < lines; i += gsize)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + i * 32));
        }
    }

    for (int i = tid; i < R * (K0 >> 2); i += blockDim.x) {
        const int r = i / (K0 >> 2), k4 = i - r * (K0 >> 2);
        const int b = min(b0 + r, p.B - 1);
        reinterpret_cast<float4*>(s_in + r * K0)[k4] = __ldg(reinterpret_cast<const float4*>(p.pooled + (size_t)b * K0) + k4);
    }
    cluster.sync();     // also: every CTA of the cluster is resident before anybody stores into its shared memory

    // ---- forward -------------------------------------------------------------------------------------------------
    for (int l = 0; l < L; ++l) {
        const float* in = l == 0 ? s_in : s_h + (l - 1) * R * H;
        dense_fwd<R, true>(cluster, p.w[l], p.b[l], in, l == 0 ? K0 : H, s_h + l * R * H, H, H, l == 0 ? K0 : H);
        cluster.sync();
    }
    dense_fwd<R, false>(cluster, p.wc, p.bc, s_h + (L - 1) * R * H, H, s_logit, Cp, C, H);
    cluster.sync();

    // ---- loss + d logits: one warp per sentence --------------------------------------------------------------------
    {
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < R && b0 + warp < p.B) {
            const int r = warp, b = b0 + r;
            const float* lg = s_logit + r * Cp;
            float m = -INFINITY;
            for (int c = lane; c < C; c += 32) m = fmaxf(m, lg[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(GPT_FULL_MASK, m, o));
            float s = 0.f;
            for (int c = lane; c < C; c += 32) s += expf(lg[c] - m);
            s = warp_sum_f(s);
            const float lse = m + logf(s);
            const int label = (int)p.labels[b];
            float l2 = 0.f;
            if (p.pooling_l2 > 0.f) {
                for (int k = lane; k < H; k += 32) { const float v = s_in[r * K0 + k]; l2 += v * v; }
                l2 = warp_sum_f(l2);
            }
            for (int c = lane; c < C; c += 32) {
                const float pr = expf(lg[c] - lse);
                const float d = (pr - (c == label ? 1.f : 0.f)) * p.inv_B;
                s_dl[r * Cp + c] = d;
                if (crank == 0) {
                    p.logits[(size_t)b * C + c] = lg[c];
                    if (p.train) p.dlogits[(size_t)b * C + c] = d;
                }
            }
            if (lane == 0 && crank == 0)
                p.loss_rows[b] = (lse - lg[label]) * p.inv_B + p.pooling_l2 * l2 * p.inv_B;
        } else if (warp < R) {
            for (int c = lane; c < C; c += 32) s_dl[warp * Cp + c] = 0.f;   // padding sentence of the last CTA
        }
    }
    if (!p.train) return;
    __syncthreads();

    // ---- backward: data gradients down to d(pooled) ----------------------------------------------------------------
    int k_lo, k_hi;
    for (int i = tid + crank * (int)blockDim.x; i < R * L * H; i += (int)blockDim.x * cs) {   // for the wgrad kernel
        const int l = i / (R * H), rem = i - l * R * H, r = rem / H, n = rem - r * H;
        if (b0 + r < p.B) p.acts[((size_t)(b0 + r) * L + l) * H + n] = s_h[i];
    }
    dense_dgrad<R>(cluster, p.wc, s_dl, Cp, s_part, s_d0, H, C, H, true, &k_lo, &k_hi);   // d post-act of last layer
    float* d_cur = s_d0;
    float* d_nxt = s_d1;
    for (int l = L - 1; l >= 0; --l) {
        for (int i = tid; i < R * H; i += blockDim.x) {           // through the ReLU; keep for dW_l
            const int r = i / H, n = i - r * H;
            const float d = s_h[(l * R + r) * H + n] > 0.f ? d_cur[i] : 0.f;
            d_cur[i] = d;
            if (crank == 0 && b0 + r < p.B) p.dacts[((size_t)(b0 + r) * L + l) * H + n] = d;
        }
        __syncthreads();
        if (l > 0) {
            dense_dgrad<R>(cluster, p.w[l], d_cur, H, s_part, d_nxt, H, H, H, true, &k_lo, &k_hi);
            float* t = d_cur; d_cur = d_nxt; d_nxt = t;
        } else {
            // d(pooled) = d_pre0 . W0 + 2 * pooling_l2 / B * [h_out, 0, 0]; s_in still holds the pooled rows
            float* s_dp = s_h;                                    // activations are no longer needed: reuse as [R][K0]
            const bool fits = L * H >= K0;                        // s_h holds L*R*H floats
            float* dst = fits ? s_dp : s_part + (size_t)R * max(kDgradPartCols, K0);   // spill area past the partials
            dense_dgrad<R>(cluster, p.w[0], d_cur, H, s_part, dst, K0, H, K0, false, &k_lo, &k_hi);
            const float c2 = 2.f * p.pooling_l2 * p.inv_B;
            const int Kc = k_hi - k_lo;                           // this rank's columns of d(pooled)
            for (int i = tid; i < R * Kc; i += blockDim.x) {
                const int r = i / Kc, k = k_lo + (i - r * Kc);
                if (b0 + r >= p.B) continue;
                float v = dst[r * K0 + k];
                if (k < H) v += c2 * s_in[r * K0 + k];
                p.dpooled[(size_t)(b0 + r) * K0 + k] = v;
            }
        }
    }
}

size_t head_smem_bytes(int R, int H, int C, int L) {
    const int K0 = 3 * H, Cp = (C + 3) & ~3;
    size_t f = (size_t)R * K0 + (size_t)L * R * H + 2 * (size_t)R * Cp + 2 * (size_t)R * H +
               (size_t)R * (K0 > kDgradPartCols ? K0 : kDgradPartCols);
    if (L * H < K0) f += (size_t)R * K0;         // separate d(pooled) staging when it does not fit over s_h
    return f * sizeof(float);
}

// ---- weight gradients ----------------------------------------------------------------------------------------------

constexpr int kWgThreads = 256, kTN = 32, kTK = 64, kTB = 32;

struct WgradParams {
    const float* pooled;        // [B, 3H]
    const float* acts;          // [B, L, H]
    const float* dacts;         // [B, L, H]
    const float* dlogits;       // [B, C]
    const float* loss_rows;     // [B]
    int B, H, C, n_mlp;
    float* dw[kMaxMlp + 1];     // per layer, classifier last
    float* db[kMaxMlp + 1];
    float* loss;                // scalar
    int tile_start[kMaxMlp + 2];
};

__global__ void __launch_bounds__(kWgThreads)
head_wgrad_kernel(const WgradParams p) {
    __shared__ float s_d[kTB][kTN + 1];
    __shared__ __align__(16) float s_x[kTB][kTK];
    __shared__ float s_red[kWgThreads / 32];
    const int L = p.n_mlp, tid = threadIdx.x;
    const int total = p.tile_start[L + 1];
    if ((int)blockIdx.x == total) {            // the extra CTA: loss = sum_b loss_rows[b], fixed order
        float s = 0.f;
        for (int b = tid; b < p.B; b += kWgThreads) s += p.loss_rows[b];
        s = warp_sum_f(s);
        if ((tid & 31) == 0) s_red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < kWgThreads / 32; ++i) t += s_red[i];
            *p.loss = t;
        }
        return;
    }
    int layer = 0;
    while (layer < L && (int)blockIdx.x >= p.tile_start[layer + 1]) ++layer;
    const int N = layer < L ? p.H : p.C;
    const int K = layer == 0 ? 3 * p.H : p.H;
    const float* dpre = layer < L ? p.dacts + (size_t)layer * p.H : p.dlogits;
    const int ld_d = layer < L ? L * p.H : p.C;
    const float* in = layer == 0 ? p.pooled : p.acts + (size_t)(layer - 1) * p.H;
    const int ld_x = layer == 0 ? 3 * p.H : L * p.H;
    const int tiles_k = (K + kTK - 1) / kTK;
    const int t = blockIdx.x - p.tile_start[layer];
    const int n_base = (t / tiles_k) * kTN, k_base = (t % tiles_k) * kTK;
    const int tn = tid >> 4, tk = tid & 15;    // 16 x 16 threads: 2 n x 4 k each
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
    float bs0 = 0.f, bs1 = 0.f;
    for (int bb = 0; bb < p.B; bb += kTB) {
        for (int i = tid; i < kTB * kTN; i += kWgThreads) {
            const int r = i / kTN, c = i - r * kTN;
            const int b = bb + r, n = n_base + c;
            s_d[r][c] = (b < p.B && n < N) ? dpre[(size_t)b * ld_d + n] : 0.f;
        }
        for (int i = tid; i < kTB * kTK; i += kWgThreads) {
            const int r = i / kTK, c = i - r * kTK;
            const int b = bb + r, k = k_base + c;
            s_x[r][c] = (b < p.B && k < K) ? in[(size_t)b * ld_x + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < kTB; ++r) {
            const float d0 = s_d[r][tn * 2], d1 = s_d[r][tn * 2 + 1];
            const float4 x = reinterpret_cast<const float4*>(&s_x[r][0])[tk];
            acc0.x += d0 * x.x; acc0.y += d0 * x.y; acc0.z += d0 * x.z; acc0.w += d0 * x.w;
            acc1.x += d1 * x.x; acc1.y += d1 * x.y; acc1.z += d1 * x.z; acc1.w += d1 * x.w;
            bs0 += d0; bs1 += d1;
        }
        __syncthreads();
    }
    float* dw = p.dw[layer];
    const float4 accs[2] = {acc0, acc1};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int n = n_base + tn * 2 + j;
        if (n >= N) continue;
        const float v[4] = {accs[j].x, accs[j].y, accs[j].z, accs[j].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k_base + tk * 4 + q;
            if (k < K) dw[(size_t)n * K + k] = v[q];
        }
        if (k_base == 0 && tk == 0) p.db[layer][n] = j == 0 ? bs0 : bs1;
    }
}

}  // namespace

extern "C" int gpt_head_fwd_bwd(const float* pooled, const int64_t* labels, const float* const* w,
                                const float* const* b, int n_mlp, const float* wc, const float* bc, int B, int H, int C,
                                float pooling_l2, int train, float* logits, float* loss_rows, float* acts, float* dacts,
                                float* dlogits, float* dpooled, void* stream) {
    GPT_CHECK_ARG(pooled && labels && w && b && wc && bc && logits && loss_rows);
    GPT_CHECK_ARG(B >= 0 && H >= 1 && C >= 1 && n_mlp >= 1 && pooling_l2 >= 0.f);
    GPT_CHECK_ARG(!train || (acts && dacts && dlogits && dpooled));
    if (n_mlp > kMaxMlp || H % 4 != 0 || C > 1024) return GPT_ERR_UNSUPPORTED;
    if (B == 0) return GPT_OK;
    HeadParams p{};
    p.pooled = pooled; p.labels = reinterpret_cast<const long long*>(labels);
    for (int l = 0; l < n_mlp; ++l) {
        GPT_CHECK_ARG(w[l] && b[l]);
        p.w[l] = w[l]; p.b[l] = b[l];
    }
    p.wc = wc; p.bc = bc;
    p.B = B; p.H = H; p.C = C; p.n_mlp = n_mlp; p.train = train;
    p.pooling_l2 = pooling_l2; p.inv_B = 1.0f / (float)B;
    p.prefetch = getenv("GPT_HEAD_NOPF") == nullptr;
    p.logits = logits; p.loss_rows = loss_rows; p.acts = acts; p.dacts = dacts; p.dlogits = dlogits; p.dpooled = dpooled;
    // sentences per CTA: 1 while a wave of CTAs fits the machine, then 2 / 4 so that weights are streamed less often
    int R = B <= 296 ? 1 : (B <= 1184 ? 2 : 4);
    if (const char* e0 = getenv("GPT_HEAD_R")) {                        // tuning knob (tools/head_bench.py)
        const int r = atoi(e0);
        if (r == 1 || r == 2 || r == 4) R = r;
    }
    const size_t smem = head_smem_bytes(R, H, C, n_mlp);
    if (smem > 200 * 1024) return GPT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (int a = R == 1 ? gpt_smem_opt_in(head_fwd_bwd_kernel<1>, smem)
                       : (R == 2 ? gpt_smem_opt_in(head_fwd_bwd_kernel<2>, smem) : gpt_smem_opt_in(head_fwd_bwd_kernel<4>, smem)))
        return a;
    const int groups = (B + R - 1) / R;
    // CTAs per sentence group: a pair of CTAs halves the number of dependent round trips to L2 per layer as long as
    // every CTA has an SM to itself (123 registers x 512 threads = one CTA per SM); wider clusters lose more to the
    // cluster barriers than they gain (tools/head_bench.py: B=50: cs 1 / 2 / 4 -> 33 / 24 / 37 us; two sentences per CTA
    // with cs = 4 -- the same 100 CTAs, a quarter of the weights each -- also lands on 24.7 us: the layer chain, not the
    // weight stream, is what bounds it)
    int cs = (R == 1 && groups * 2 <= 148) ? 2 : 1;
    if (const char* e2 = getenv("GPT_HEAD_CS")) cs = atoi(e2) > 0 ? atoi(e2) : cs;   // tuning knob (tools/head_bench.py)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(groups * cs));
    cfg.blockDim = dim3(kHeadThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_gpt_pdl ? 2 : 1;
    if (R == 1) e = cudaLaunchKernelEx(&cfg, head_fwd_bwd_kernel<1>, p);
    else if (R == 2) e = cudaLaunchKernelEx(&cfg, head_fwd_bwd_kernel<2>, p);
    else e = cudaLaunchKernelEx(&cfg, head_fwd_bwd_kernel<4>, p);
    ++g_gpt_launches;
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    return e == cudaSuccess ? GPT_OK : (int)e;
}

extern "C" int gpt_head_wgrad(const float* pooled, const float* acts, const float* dacts, const float* dlogits,
                              const float* loss_rows, int B, int H, int C, int n_mlp, float* const* dw,
                              float* const* db, float* dwc, float* dbc, float* loss, void* stream) {
    GPT_CHECK_ARG(pooled && acts && dacts && dlogits && loss_rows && dw && db && dwc && dbc && loss);
    GPT_CHECK_ARG(B >= 1 && H >= 1 && C >= 1 && n_mlp >= 1);
    if (n_mlp > kMaxMlp) return GPT_ERR_UNSUPPORTED;
    WgradParams p{};
    p.pooled = pooled; p.acts = acts; p.dacts = dacts; p.dlogits = dlogits; p.loss_rows = loss_rows;
    p.B = B; p.H = H; p.C = C; p.n_mlp = n_mlp; p.loss = loss;
    int tiles = 0;
    for (int l = 0; l <= n_mlp; ++l) {
        const int N = l < n_mlp ? H : C, K = l == 0 ? 3 * H : H;
        p.tile_start[l] = tiles;
        tiles += ((N + kTN - 1) / kTN) * ((K + kTK - 1) / kTK);
        if (l < n_mlp) {
            GPT_CHECK_ARG(dw[l] && db[l]);
            p.dw[l] = dw[l]; p.db[l] = db[l];
        } else {
            p.dw[l] = dwc; p.db[l] = dbc;
        }
    }
    p.tile_start[n_mlp + 1] = tiles;
    head_wgrad_kernel<<<tiles + 1, kWgThreads, 0, (cudaStream_t)stream>>>(p);
    return gpt_launch_status();
}
