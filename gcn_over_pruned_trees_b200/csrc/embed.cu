// K5 -- input embedding stage of GCN.forward and its row-sparse gradient / optimiser path.
//
// Forward replaces  cat[emb(words), pos_emb(pos), ner_emb(ner)] -> in_drop   (/root/reference/model/gcn.py:235-247)
// -- three gather kernels, a concat and a dropout in the reference -- with one kernel that writes the layer-0
// input row [emb_dim + pos_dim + ner_dim] once, with the dropout mask drawn in-kernel (Philox, same {seed, step}
// device state as the GCN layers).
//
// Backward: the reference's word-embedding gradient is a dense [V, emb_dim] tensor that is zero outside the <= B*T
// rows the batch touches (60 MB at V = 50 000 for ~1 000 live rows); zeroing, clipping and updating it densely
// dominates a TACRED-size step on a GPU.  Here dX rows are scattered into the tables with atomics (rows of tokens
// that are not observable have exactly zero gradient and are skipped), and -- for the plain-SGD configuration the
// reference ships (train_gcn.sh:4) -- three small kernels finish the step touching only the live rows:
//   embed_bwd      : G[w] += dX[t] * mask ; owner[w] = min token index with that word (dedup for the passes below)
//   rows_sqnorm    : sq += sum over owned rows of |G[w]|^2          (the embedding's share of the global grad norm)
//   rows_sgd       : W[w] -= lr * clip * G[w] ; G[w] = 0 ; owner[w] = INT_MAX      (G stays all-zero between steps)
// which is arithmetic-identical to clip_grad_norm_ + SGD over the dense tensor (train.py:224-227).
#include "gpt_common.cuh"
#include <cstdlib>

namespace {

constexpr int kEmbThreads = 128;

struct EmbParams {
    const long long* words;
    const long long* pos;
    const long long* ner;      // may be null (SemEval / ner_dim == 0)
    const float* emb_w;        // [V, E]
    const float* pos_w;        // [P, Dp] or null
    const float* ner_w;        // [N, Dn] or null
    int n_rows, V, E, Dp, Dn;
    unsigned thresh16, subseq;
    float drop_scale;
    const unsigned long long* rng;
};

// one CTA per token row; thread c handles column group c*8 .. c*8+7 (one Philox call per thread)
__global__ void __launch_bounds__(kEmbThreads)
embed_fwd_kernel(const EmbParams p, float* __restrict__ x) {
    GPT_PDL_ENTER();
    const int row = blockIdx.x;
    const int D = p.E + p.Dp + p.Dn;
    const long long w = p.words[row];
    const long long ps = p.pos_w ? p.pos[row] : 0;
    const long long nr = p.ner_w ? p.ner[row] : 0;
    const bool drop = p.thresh16 > 0;
    unsigned long long seed = 0, step = 0;
    if (drop) { seed = p.rng[0]; step = p.rng[1]; }
    float* xr = x + (size_t)row * D;
    for (int g = threadIdx.x; g * 8 < D; g += kEmbThreads) {
        Philox4 q{0, 0, 0, 0};
        if (drop)
            q = philox4x32((uint32_t)g | (p.subseq << 20), (uint32_t)row, 0x454d4245u, (uint32_t)step, (uint32_t)seed,
                           (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
        const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            if (c >= D) break;
            float v;
            if (c < p.E) v = p.emb_w[(size_t)w * p.E + c];
            else if (c < p.E + p.Dp) v = p.pos_w[(size_t)ps * p.Dp + (c - p.E)];
            else v = p.ner_w[(size_t)nr * p.Dn + (c - p.E - p.Dp)];
            if (drop) {
                const uint32_t bits = (r4[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                v = bits >= p.thresh16 ? v * p.drop_scale : 0.f;
            }
            xr[c] = v;
        }
    }
}

// The same rows for LARGE batches (>= 64 K token rows): one CTA per row means 2 M CTAs of 45 busy threads at the large
// shape -- 3.2 ms, bound by the rate at which CTAs can be launched, for 3 GB of output.  Here a resident grid walks
// (row, 8-column group) items with every thread busy; same Philox stream, bit-identical output.
__device__ __forceinline__ void embed_rows_body(const EmbParams& p, float* __restrict__ x, unsigned block, unsigned n_blocks) {
    const int D = p.E + p.Dp + p.Dn, ng = (D + 7) >> 3;
    const bool drop = p.thresh16 > 0;
    unsigned long long seed = 0, step = 0;
    if (drop) { seed = p.rng[0]; step = p.rng[1]; }
    const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const long long total = (long long)p.n_rows * ng;
    for (long long it = (long long)block * blockDim.x + threadIdx.x; it < total; it += (long long)n_blocks * blockDim.x) {
        const int row = (int)(it / ng), g = (int)(it - (long long)row * ng);
        const long long w = p.words[row];
        const long long ps = p.pos_w ? p.pos[row] : 0;
        const long long nr = p.ner_w ? p.ner[row] : 0;
        Philox4 q{0, 0, 0, 0};
        if (drop)
            q = philox4x32((uint32_t)g | (p.subseq << 20), (uint32_t)row, 0x454d4245u, (uint32_t)step, (uint32_t)seed,
                           (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
        const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            float u = 0.f;
            if (c < p.E) u = p.emb_w[(size_t)w * p.E + c];
            else if (c < p.E + p.Dp) u = p.pos_w[(size_t)ps * p.Dp + (c - p.E)];
            else if (c < D) u = p.ner_w[(size_t)nr * p.Dn + (c - p.E - p.Dp)];
            if (drop) {
                const uint32_t bits = (r4[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                u = bits >= p.thresh16 ? u * p.drop_scale : 0.f;
            }
            v[k] = u;
        }
        float* xr = x + (size_t)row * D + g * 8;
        if (vec && g * 8 + 7 < D) {
            reinterpret_cast<float4*>(xr)[0] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(xr)[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (g * 8 + k < D) xr[k] = v[k];
        }
    }
}

__global__ void __launch_bounds__(256)
embed_fwd_rows_kernel(const EmbParams p, float* __restrict__ x) {
    GPT_PDL_ENTER();
    embed_rows_body(p, x, blockIdx.x, gridDim.x);
}

// The FRONT of a training step in one launch: the embedding rows (blocks [0, embed_blocks)) and, in the blocks behind them,
// the per-step operand preparation of every GCN layer's weight for the 3xTF32 projections -- w -> [w_hi | w_lo | w^T_hi |
// w^T_lo], hi = round_tf32, lo = w - hi, 32 x 32 tiles through shared memory (what gpt_weight_prep_tf32x3_batch does in a
// launch of its own).  As separate root nodes of the captured step the two kernels start ~6 us apart and the first
// projection waits for the later one across streams: it began at 17.7 us of the replay; behind this kernel, 12 us.
struct FrontPrep {
    const float* w[8];
    float* ws[8];
    int N[8], K[8];
    int tile0[9];            // first tile of layer l in the flat tile index; tile0[n_layers] = number of tiles
    int n_layers;
};

__device__ __forceinline__ float front_tf32_hi(float v) {     // round to nearest TF32 (ties away), as tc::tf32_hi
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

__global__ void __launch_bounds__(256)
embed_fwd_rows_prep_kernel(const EmbParams p, float* __restrict__ x, const FrontPrep f, int embed_blocks) {
    GPT_PDL_ENTER();
    if ((int)blockIdx.x < embed_blocks) {
        embed_rows_body(p, x, blockIdx.x, (unsigned)embed_blocks);
        return;
    }
    __shared__ float th[32][33], tl[32][33];
    const int tile = (int)blockIdx.x - embed_blocks;
    int l = 0;
    while (l + 1 < f.n_layers && tile >= f.tile0[l + 1]) ++l;
    const float* __restrict__ w = f.w[l];
    float* __restrict__ ws = f.ws[l];
    const int N = f.N[l], K = f.K[l];
    const int tiles_x = (K + 31) / 32, t = tile - f.tile0[l];
    const int bx = t % tiles_x, by = t / tiles_x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t nk = (size_t)N * K;
    const int xk = bx * 32 + tx, y0 = by * 32;
    for (int j = ty; j < 32; j += 8) {
        if (xk < K && y0 + j < N) {
            const size_t i = (size_t)(y0 + j) * K + xk;
            const float v = w[i], h = front_tf32_hi(v);
            ws[i] = h;
            ws[nk + i] = v - h;
            th[j][tx] = h;
            tl[j][tx] = v - h;
        }
    }
    __syncthreads();
    const int ox = by * 32 + tx, oy0 = bx * 32;
    for (int j = ty; j < 32; j += 8) {
        if (ox < N && oy0 + j < K) {
            const size_t o = (size_t)(oy0 + j) * N + ox;
            ws[2 * nk + o] = th[tx][j];
            ws[3 * nk + o] = tl[tx][j];
        }
    }
}

// scatter dX (after the same dropout mask) into the tables; word rows additionally record their first token.
// MULTI = false: one CTA per token row (TACRED-sized batches: thousands of short CTAs).  MULTI = true (>= 64 K rows):
// a CTA walks many rows and first collects the gradients of the small tables -- 47 POS and 15 NER rows, i.e. millions of
// atomics onto ~60 cache lines at the large shape -- in shared memory (ids < kSmallIds; larger ids go to memory
// directly) and flushes each touched entry once.
constexpr int kSmallIds = 64;

template <bool MULTI>
__global__ void __launch_bounds__(kEmbThreads)
embed_bwd_kernel(const EmbParams p, const float* __restrict__ dx, const unsigned char* __restrict__ flags,
                 float* __restrict__ g_emb, float* __restrict__ g_pos, float* __restrict__ g_ner,
                 int* __restrict__ owner, int topn) {
    GPT_PDL_ENTER();
    extern __shared__ float s_small[];                   // MULTI: [kSmallIds][Dp] then [kSmallIds][Dn]
    float* s_pos = s_small;
    float* s_ner = s_small + kSmallIds * p.Dp;
    if (MULTI) {
        for (int i = threadIdx.x; i < kSmallIds * (p.Dp + p.Dn); i += kEmbThreads) s_small[i] = 0.f;
        __syncthreads();
    }
    const int D = p.E + p.Dp + p.Dn;
    const bool drop = p.thresh16 > 0;
    unsigned long long seed = 0, step = 0;
    if (drop) { seed = p.rng[0]; step = p.rng[1]; }
    const bool vec_rows = (p.E & 3) == 0 && (reinterpret_cast<uintptr_t>(g_emb) & 15) == 0;   // 16-byte aligned quads
    for (int row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        if (flags != nullptr && flags[row] == 0) continue;  // unobservable token: its gradient row is exactly zero
        const long long w = p.words[row];
        const long long ps = p.pos_w ? p.pos[row] : 0;
        const long long nr = p.ner_w ? p.ner[row] : 0;
        const bool word_live = g_emb != nullptr && w != 0 && w < topn;  // padding_idx = 0; rows >= topn are frozen
        if (threadIdx.x == 0 && word_live && owner != nullptr) atomicMin(owner + w, row);
        const float* dr = dx + (size_t)row * D;
        for (int g = threadIdx.x; g * 8 < D; g += kEmbThreads) {
            Philox4 q{0, 0, 0, 0};
            if (drop)
                q = philox4x32((uint32_t)g | (p.subseq << 20), (uint32_t)row, 0x454d4245u, (uint32_t)step,
                               (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
            const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = g * 8 + k;
                v[k] = c < D ? dr[c] : 0.f;
                if (drop) {
                    const uint32_t bits = (r4[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                    v[k] = bits >= p.thresh16 ? v[k] * p.drop_scale : 0.f;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {                   // two quads of columns
                const int c = g * 8 + 4 * h;
                if (c >= D) break;
                if (vec_rows && c + 3 < p.E) {              // whole quad in the word row: one 16-byte reduction
                    if (word_live && (v[4 * h] != 0.f || v[4 * h + 1] != 0.f || v[4 * h + 2] != 0.f || v[4 * h + 3] != 0.f))
#ifdef GPT_HOST_EMULATION   // tests/emu
                        for (int k = 0; k < 4; ++k) atomicAdd(g_emb + (size_t)w * p.E + c + k, v[4 * h + k]);
#else
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g_emb + (size_t)w * p.E + c),
                                     "f"(v[4 * h]), "f"(v[4 * h + 1]), "f"(v[4 * h + 2]), "f"(v[4 * h + 3]) : "memory");
#endif
                    continue;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int cc = c + k;
                    const float u = v[4 * h + k];
                    if (cc >= D || u == 0.f) continue;
                    if (cc < p.E) {
                        if (word_live) atomicAdd(g_emb + (size_t)w * p.E + cc, u);
                    } else if (cc < p.E + p.Dp) {
                        if (g_pos) {
                            if (MULTI && ps < kSmallIds) atomicAdd(s_pos + (int)ps * p.Dp + (cc - p.E), u);
                            else atomicAdd(g_pos + (size_t)ps * p.Dp + (cc - p.E), u);
                        }
                    } else if (g_ner) {
                        if (MULTI && nr < kSmallIds) atomicAdd(s_ner + (int)nr * p.Dn + (cc - p.E - p.Dp), u);
                        else atomicAdd(g_ner + (size_t)nr * p.Dn + (cc - p.E - p.Dp), u);
                    }
                }
            }
        }
    }
    if (MULTI) {
        __syncthreads();
        for (int i = threadIdx.x; i < kSmallIds * p.Dp; i += kEmbThreads)
            if (g_pos && s_pos[i] != 0.f) atomicAdd(g_pos + i, s_pos[i]);        // same [id][column] layout as the table
        for (int i = threadIdx.x; i < kSmallIds * p.Dn; i += kEmbThreads)
            if (g_ner && s_ner[i] != 0.f) atomicAdd(g_ner + i, s_ner[i]);
    }
}

// ---- word-embedding gradient of LARGE batches: group the token rows by word, then one plain sum per word -----------------
// At the large shape (2 M tokens over a 50 k vocabulary) the scatter above is 157 M 16-byte reductions into a 60 MB
// table: 10.2 ms, bound by L2 reduction throughput, not by the 3 GB of dX it reads.  Grouped: a counting sort of the live
// token rows by word id (count -> scan -> fill; integer atomics on 50 k counters), then one warp per word adds its ~40
// rows of dX in registers (dropout mask re-derived per row, as above) and touches G[w] once: the kernel streams dX at HBM
// speed and issues no floating-point atomic at all.  (The order of a word's rows inside its group depends on the fill's
// atomics, like the order of the reductions it replaces.)
__global__ void __launch_bounds__(256)
emb_group_zero_kernel(int* __restrict__ count, int n) {
    GPT_PDL_ENTER();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) count[i] = 0;
}

__global__ void __launch_bounds__(256)
emb_group_count_kernel(const long long* __restrict__ words, const unsigned char* __restrict__ flags, int n_rows, int topn,
                       int* __restrict__ count) {
    GPT_PDL_ENTER();
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
        if (flags != nullptr && flags[row] == 0) continue;
        const long long w = words[row];
        if (w != 0 && w < topn) atomicAdd(count + w, 1);
    }
}

// start[w] = number of live rows of words < w (start[V] = total); cursor[w] = start[w].  One CTA.
__global__ void __launch_bounds__(1024)
emb_group_scan_kernel(const int* __restrict__ count, int V, int* __restrict__ start, int* __restrict__ cursor) {
    GPT_PDL_ENTER();
    __shared__ int s_sum[1024];
    const int per = (V + 1023) / 1024, lo = threadIdx.x * per, hi = min(V, lo + per);
    int t = 0;
    for (int i = lo; i < hi; ++i) t += count[i];
    s_sum[threadIdx.x] = t;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {            // inclusive scan of the per-thread totals
        const int v = threadIdx.x >= o ? s_sum[threadIdx.x - o] : 0;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_sum[threadIdx.x] - t;
    for (int i = lo; i < hi; ++i) {
        start[i] = run;
        cursor[i] = run;
        run += count[i];
    }
    if (threadIdx.x == 1023) start[V] = s_sum[1023];
}

__global__ void __launch_bounds__(256)
emb_group_fill_kernel(const long long* __restrict__ words, const unsigned char* __restrict__ flags, int n_rows, int topn,
                      int* __restrict__ cursor, int* __restrict__ rows_sorted) {
    GPT_PDL_ENTER();
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
        if (flags != nullptr && flags[row] == 0) continue;
        const long long w = words[row];
        if (w != 0 && w < topn) rows_sorted[atomicAdd(cursor + w, 1)] = row;
    }
}

// one warp per word: G[w] += sum over the word's rows of mask(row) * dX[row, :E];  owner[w] = its smallest row.
// Lane l serves the 8-column groups l and l + 32 (E <= 512); needs E % 4 == 0 and (E + Dp + Dn) % 4 == 0 (16-byte loads).
__global__ void __launch_bounds__(256)
emb_group_sum_kernel(const EmbParams p, const float* __restrict__ dx, const int* __restrict__ start,
                     const int* __restrict__ rows_sorted, float* __restrict__ g_emb, int* __restrict__ owner) {
    GPT_PDL_ENTER();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int D = p.E + p.Dp + p.Dn;
    const bool drop = p.thresh16 > 0;
    unsigned long long seed = 0, step = 0;
    if (drop) { seed = p.rng[0]; step = p.rng[1]; }
    const int ng = (p.E + 7) >> 3;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < p.V; w += warps) {
        const int s0 = start[w], n = start[w + 1] - s0;
        if (n == 0) continue;                                           // warp-uniform
        float acc[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[u][k] = 0.f;
        int min_row = 0x7fffffff;
        for (int i = 0; i < n; ++i) {
            const int row = rows_sorted[s0 + i];
            min_row = min(min_row, row);
            const float* dr = dx + (size_t)row * D;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int g = lane + 32 * u;
                if (g >= ng) break;
                const int c = g * 8;
                const float4 a = *reinterpret_cast<const float4*>(dr + c);
                const float4 b = c + 4 < p.E ? *reinterpret_cast<const float4*>(dr + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                if (drop) {
                    const Philox4 q = philox4x32((uint32_t)g | (p.subseq << 20), (uint32_t)row, 0x454d4245u, (uint32_t)step,
                                                 (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
                    const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t bits = (r4[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                        v[k] = bits >= p.thresh16 ? v[k] * p.drop_scale : 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[u][k] += v[k];
            }
        }
        float* gr = g_emb + (size_t)w * p.E;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int g = lane + 32 * u;
            if (g >= ng) break;
            const int c = g * 8;
            float4* q0 = reinterpret_cast<float4*>(gr + c);
            float4 t = *q0;
            t.x += acc[u][0]; t.y += acc[u][1]; t.z += acc[u][2]; t.w += acc[u][3];
            *q0 = t;
            if (c + 4 < p.E) {
                float4* q1 = reinterpret_cast<float4*>(gr + c + 4);
                float4 t1 = *q1;
                t1.x += acc[u][4]; t1.y += acc[u][5]; t1.z += acc[u][6]; t1.w += acc[u][7];
                *q1 = t1;
            }
        }
        if (lane == 0 && owner != nullptr) owner[w] = min(owner[w], min_row);
    }
}

// The POS / NER columns of dX (the last Dp + Dn of every row) for large batches: a thread serves one 8-column group of
// one row, so every thread of the CTA is busy (the scatter kernel's row loop leaves 120 of its 128 threads idle on these
// 60 columns); gradients meet in shared memory (ids < kSmallIds) and every touched entry is flushed once per CTA.
__global__ void __launch_bounds__(256)
emb_small_tables_kernel(const EmbParams p, const float* __restrict__ dx, const unsigned char* __restrict__ flags,
                        float* __restrict__ g_pos, float* __restrict__ g_ner) {
    GPT_PDL_ENTER();
    extern __shared__ float s_small[];                   // [kSmallIds][Dp] then [kSmallIds][Dn]
    float* s_pos = s_small;
    float* s_ner = s_small + kSmallIds * p.Dp;
    for (int i = threadIdx.x; i < kSmallIds * (p.Dp + p.Dn); i += blockDim.x) s_small[i] = 0.f;
    __syncthreads();
    const int D = p.E + p.Dp + p.Dn;
    const int g0 = p.E >> 3, tpr = ((D + 7) >> 3) - g0;              // 8-column groups that hold small-table columns
    const int rows_per_pass = blockDim.x / tpr;
    const int lr = threadIdx.x / tpr, g = g0 + threadIdx.x % tpr;
    const bool drop = p.thresh16 > 0;
    unsigned long long seed = 0, step = 0;
    if (drop) { seed = p.rng[0]; step = p.rng[1]; }
    if (lr < rows_per_pass) {
        for (long long row = (long long)blockIdx.x * rows_per_pass + lr; row < p.n_rows;
             row += (long long)gridDim.x * rows_per_pass) {
            if (flags != nullptr && flags[row] == 0) continue;
            const long long ps = p.pos_w ? p.pos[row] : 0;
            const long long nr = p.ner_w ? p.ner[row] : 0;
            const float* dr = dx + (size_t)row * D;
            Philox4 q{0, 0, 0, 0};
            if (drop)
                q = philox4x32((uint32_t)g | (p.subseq << 20), (uint32_t)row, 0x454d4245u, (uint32_t)step, (uint32_t)seed,
                               (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
            const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int cc = g * 8 + k;
                if (cc < p.E || cc >= D) continue;
                float u = dr[cc];
                if (drop) {
                    const uint32_t bits = (r4[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                    u = bits >= p.thresh16 ? u * p.drop_scale : 0.f;
                }
                if (u == 0.f) continue;
                if (cc < p.E + p.Dp) {
                    if (g_pos) {
                        if (ps < kSmallIds) atomicAdd(s_pos + (int)ps * p.Dp + (cc - p.E), u);
                        else atomicAdd(g_pos + (size_t)ps * p.Dp + (cc - p.E), u);
                    }
                } else if (g_ner) {
                    if (nr < kSmallIds) atomicAdd(s_ner + (int)nr * p.Dn + (cc - p.E - p.Dp), u);
                    else atomicAdd(g_ner + (size_t)nr * p.Dn + (cc - p.E - p.Dp), u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSmallIds * p.Dp; i += blockDim.x)
        if (g_pos && s_pos[i] != 0.f) atomicAdd(g_pos + i, s_pos[i]);
    for (int i = threadIdx.x; i < kSmallIds * p.Dn; i += blockDim.x)
        if (g_ner && s_ner[i] != 0.f) atomicAdd(g_ner + i, s_ner[i]);
}

__global__ void __launch_bounds__(kEmbThreads)
rows_sqnorm_kernel(const long long* __restrict__ words, const int* __restrict__ owner, const float* __restrict__ g,
                   int n_rows, int E, int topn, float* __restrict__ sq) {
    const int row = blockIdx.x;
    const long long w = words[row];
    if (w == 0 || w >= topn || owner[w] != row) return;  // every live word row is counted once, by its first token
    float s = 0.f;
    for (int c = threadIdx.x; c < E; c += kEmbThreads) { const float v = g[(size_t)w * E + c]; s += v * v; }
    s = warp_sum_f(s);
    __shared__ float part[kEmbThreads / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < kEmbThreads / 32; ++i) t += part[i];
        atomicAdd(sq, t);
    }
}

// W[w] -= lr * min(1, max_norm / (sqrt(total_sq) + 1e-6)) * G[w]; G[w] = 0; owner[w] = INT_MAX
__global__ void __launch_bounds__(kEmbThreads)
rows_sgd_kernel(const long long* __restrict__ words, int* __restrict__ owner, float* __restrict__ g,
                float* __restrict__ w_tab, int n_rows, int E, int topn, const float* __restrict__ total_sq,
                float max_norm, float lr) {
    const int row = blockIdx.x;
    const long long w = words[row];
    if (w == 0 || w >= topn || owner[w] != row) return;
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*total_sq) + 1e-6f));  // clip_grad_norm_ semantics
    const float a = lr * coef;
    for (int c = threadIdx.x; c < E; c += kEmbThreads) {
        const size_t i = (size_t)w * E + c;
        w_tab[i] -= a * g[i];
        g[i] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) owner[w] = 0x7fffffff;
}

int fill_params(EmbParams& p, const int64_t* words, const int64_t* pos, const int64_t* ner, const float* emb_w,
                const float* pos_w, const float* ner_w, int n_rows, int V, int E, int Dp, int Dn, float drop_p,
                const uint64_t* rng, uint32_t subseq) {
    if (!(words && emb_w) || n_rows < 0 || V < 1 || E < 1 || Dp < 0 || Dn < 0) return GPT_ERR_BAD_ARG;
    if ((Dp > 0) != (pos_w != nullptr && pos != nullptr)) return GPT_ERR_BAD_ARG;
    if ((Dn > 0) != (ner_w != nullptr && ner != nullptr)) return GPT_ERR_BAD_ARG;
    if (!(drop_p >= 0.f && drop_p < 1.f) || (drop_p > 0.f && rng == nullptr)) return GPT_ERR_BAD_ARG;
    p.words = reinterpret_cast<const long long*>(words);
    p.pos = reinterpret_cast<const long long*>(pos);
    p.ner = reinterpret_cast<const long long*>(ner);
    p.emb_w = emb_w; p.pos_w = pos_w; p.ner_w = ner_w;
    p.n_rows = n_rows; p.V = V; p.E = E; p.Dp = Dp; p.Dn = Dn;
    unsigned th = (unsigned)(drop_p * 65536.0f + 0.5f);
    p.thresh16 = th > 65535u ? 65535u : th;
    p.drop_scale = p.thresh16 > 0 ? 65536.0f / (65536.0f - (float)p.thresh16) : 1.0f;
    p.subseq = subseq & 0xfffu;
    p.rng = reinterpret_cast<const unsigned long long*>(rng);
    return GPT_OK;
}

}  // namespace

extern "C" int gpt_embed_fwd(const int64_t* words, const int64_t* pos, const int64_t* ner, const float* emb_w,
                             const float* pos_w, const float* ner_w, float* x, int n_rows, int V, int E, int Dp, int Dn,
                             float drop_p, const uint64_t* rng_state, uint32_t subseq, void* stream) {
    EmbParams p{};
    int rc = fill_params(p, words, pos, ner, emb_w, pos_w, ner_w, n_rows, V, E, Dp, Dn, drop_p, rng_state, subseq);
    if (rc != GPT_OK || x == nullptr) return rc != GPT_OK ? rc : GPT_ERR_BAD_ARG;
    if (n_rows == 0) return GPT_OK;
    static const int rows_min = [] { const char* e = getenv("GPT_EMBED_ROWS_MIN"); return e ? atoi(e) : 65536; }();
    if (n_rows >= rows_min) {
        const long long items = (long long)n_rows * ((E + Dp + Dn + 7) / 8);
        long long blocks = (items + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        gpt_launch(embed_fwd_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, p, x);
        return gpt_launch_status();
    }
    gpt_launch(embed_fwd_kernel, dim3(n_rows), dim3(kEmbThreads), 0, (cudaStream_t)stream, p, x);
    return gpt_launch_status();
}

// gpt_embed_fwd (resident-grid form) and gpt_weight_prep_tf32x3_batch in ONE launch (see embed_fwd_rows_prep_kernel)
extern "C" int gpt_embed_fwd_prep(const int64_t* words, const int64_t* pos, const int64_t* ner, const float* emb_w,
                                  const float* pos_w, const float* ner_w, float* x, int n_rows, int V, int E, int Dp, int Dn,
                                  float drop_p, const uint64_t* rng_state, uint32_t subseq, const float* const* w,
                                  float* const* ws, const int* wN, const int* wK, int n_layers, void* stream) {
    EmbParams p{};
    int rc = fill_params(p, words, pos, ner, emb_w, pos_w, ner_w, n_rows, V, E, Dp, Dn, drop_p, rng_state, subseq);
    if (rc != GPT_OK || x == nullptr) return rc != GPT_OK ? rc : GPT_ERR_BAD_ARG;
    GPT_CHECK_ARG(w && ws && wN && wK && n_layers >= 1 && n_layers <= 8);
    FrontPrep f{};
    f.n_layers = n_layers;
    int tiles = 0;
    for (int l = 0; l < n_layers; ++l) {
        GPT_CHECK_ARG(w[l] && ws[l] && wN[l] >= 1 && wK[l] >= 1);
        f.w[l] = w[l]; f.ws[l] = ws[l]; f.N[l] = wN[l]; f.K[l] = wK[l];
        f.tile0[l] = tiles;
        tiles += ((wK[l] + 31) / 32) * ((wN[l] + 31) / 32);
    }
    f.tile0[n_layers] = tiles;
    const long long items = (long long)n_rows * ((E + Dp + Dn + 7) / 8);
    long long blocks = (items + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    gpt_launch(embed_fwd_rows_prep_kernel, dim3((unsigned)(blocks + tiles)), dim3(256), 0, (cudaStream_t)stream, p, x, f,
               (int)blocks);
    return gpt_launch_status();
}

extern "C" int gpt_embed_bwd(const float* dx, const uint8_t* flags, const int64_t* words, const int64_t* pos,
                             const int64_t* ner, float* g_emb, float* g_pos, float* g_ner, int32_t* owner, int n_rows,
                             int V, int E, int Dp, int Dn, int topn, float drop_p, const uint64_t* rng_state,
                             uint32_t subseq, void* stream) {
    EmbParams p{};
    // the tables themselves are not read in the backward; reuse the checker with dummy non-null table pointers
    int rc = fill_params(p, words, pos, ner, dx, Dp > 0 ? dx : nullptr, Dn > 0 ? dx : nullptr, n_rows, V, E, Dp, Dn,
                         drop_p, rng_state, subseq);
    if (rc != GPT_OK || dx == nullptr) return rc != GPT_OK ? rc : GPT_ERR_BAD_ARG;
    if (n_rows == 0) return GPT_OK;
    const cudaStream_t st = (cudaStream_t)stream;
    const int tn = topn < V ? topn : V;
    if (n_rows >= 65536) {      // many rows: a few resident CTAs walk them, small-table gradients meet in shared memory
        const size_t smem = (size_t)kSmallIds * (Dp + Dn) * sizeof(float);
        if (smem <= 48 * 1024) {
            gpt_launch(embed_bwd_kernel<true>, dim3(148 * 8), dim3(kEmbThreads), smem, st, p, dx, flags, g_emb, g_pos,
                       g_ner, owner, tn);
            return gpt_launch_status();
        }
    }
    gpt_launch(embed_bwd_kernel<false>, dim3(n_rows), dim3(kEmbThreads), 0, st, p, dx, flags, g_emb, g_pos, g_ner, owner,
               tn);
    return gpt_launch_status();
}

extern "C" long long gpt_embed_bwd_grouped_workspace(int n_rows, int V) {      // int32 entries
    return 3ll * V + 1 + n_rows;
}

// The same gradients as gpt_embed_bwd for large batches, without floating-point atomics on the word table: the live rows
// are grouped by word (workspace: gpt_embed_bwd_grouped_workspace(n_rows, V) int32) and each word's rows are summed by one
// warp; the POS / NER columns go through the shared-memory path of gpt_embed_bwd.  GPT_ERR_UNSUPPORTED when the row
// widths do not allow 16-byte loads (the caller falls back to gpt_embed_bwd).
extern "C" int gpt_embed_bwd_grouped(const float* dx, const uint8_t* flags, const int64_t* words, const int64_t* pos,
                                     const int64_t* ner, float* g_emb, float* g_pos, float* g_ner, int32_t* owner,
                                     int n_rows, int V, int E, int Dp, int Dn, int topn, float drop_p,
                                     const uint64_t* rng_state, uint32_t subseq, int32_t* workspace, void* stream) {
    EmbParams p{};
    int rc = fill_params(p, words, pos, ner, dx, Dp > 0 ? dx : nullptr, Dn > 0 ? dx : nullptr, n_rows, V, E, Dp, Dn,
                         drop_p, rng_state, subseq);
    if (rc != GPT_OK || dx == nullptr || g_emb == nullptr || workspace == nullptr) return rc != GPT_OK ? rc : GPT_ERR_BAD_ARG;
    if (n_rows == 0) return GPT_OK;
    const int D = E + Dp + Dn;
    if (E % 4 != 0 || D % 4 != 0 || E > 512 || (reinterpret_cast<uintptr_t>(dx) & 15) || (reinterpret_cast<uintptr_t>(g_emb) & 15))
        return GPT_ERR_UNSUPPORTED;
    const size_t smem = (size_t)kSmallIds * (Dp + Dn) * sizeof(float);
    if (smem > 48 * 1024) return GPT_ERR_UNSUPPORTED;
    const cudaStream_t st = (cudaStream_t)stream;
    const int tn = topn < V ? topn : V;
    int* count = workspace;
    int* start = workspace + V;             // [V + 1]
    int* cursor = start + V + 1;            // [V]
    int* rows_sorted = cursor + V;          // [n_rows]
    const long long* w64 = reinterpret_cast<const long long*>(words);
    gpt_launch(emb_group_zero_kernel, dim3(64), dim3(256), 0, st, count, V);
    if ((rc = gpt_launch_status()) != GPT_OK) return rc;
    gpt_launch(emb_group_count_kernel, dim3(148 * 8), dim3(256), 0, st, w64, flags, n_rows, tn, count);
    if ((rc = gpt_launch_status()) != GPT_OK) return rc;
    gpt_launch(emb_group_scan_kernel, dim3(1), dim3(1024), 0, st, (const int*)count, V, start, cursor);
    if ((rc = gpt_launch_status()) != GPT_OK) return rc;
    gpt_launch(emb_group_fill_kernel, dim3(148 * 8), dim3(256), 0, st, w64, flags, n_rows, tn, cursor, rows_sorted);
    if ((rc = gpt_launch_status()) != GPT_OK) return rc;
    gpt_launch(emb_group_sum_kernel, dim3(148 * 8), dim3(256), 0, st, p, dx, (const int*)start, (const int*)rows_sorted,
               g_emb, owner);
    if ((rc = gpt_launch_status()) != GPT_OK) return rc;
    if ((g_pos != nullptr || g_ner != nullptr) && Dp + Dn > 0) {     // the POS / NER columns: every thread busy
        gpt_launch(emb_small_tables_kernel, dim3(148 * 4), dim3(256), smem, st, p, dx, flags, g_pos, g_ner);
        rc = gpt_launch_status();
    }
    return rc;
}

extern "C" int gpt_embed_rows_sqnorm(const int64_t* words, const int32_t* owner, const float* g_emb, int n_rows, int E,
                                     int topn, float* sq, void* stream) {
    GPT_CHECK_ARG(words && owner && g_emb && sq && n_rows >= 0 && E >= 1);
    if (n_rows == 0) return GPT_OK;
#ifdef GPT_HOST_EMULATION   // tests/emu: g++ has no <<<>>>
    gpt_launch(rows_sqnorm_kernel, dim3(n_rows), dim3(kEmbThreads), 0, (cudaStream_t)stream,
               reinterpret_cast<const long long*>(words), owner, g_emb, n_rows, E, topn, sq);
#else
    rows_sqnorm_kernel<<<n_rows, kEmbThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(words),
                                                                         owner, g_emb, n_rows, E, topn, sq);
#endif
    return gpt_launch_status();
}

extern "C" int gpt_embed_rows_sgd(const int64_t* words, int32_t* owner, float* g_emb, float* emb_w, int n_rows, int E,
                                  int topn, const float* total_sq, float max_norm, float lr, void* stream) {
    GPT_CHECK_ARG(words && owner && g_emb && emb_w && n_rows >= 0 && E >= 1);
    GPT_CHECK_ARG(max_norm <= 0.f || total_sq != nullptr);
    if (n_rows == 0) return GPT_OK;
#ifdef GPT_HOST_EMULATION
    gpt_launch(rows_sgd_kernel, dim3(n_rows), dim3(kEmbThreads), 0, (cudaStream_t)stream,
               reinterpret_cast<const long long*>(words), owner, g_emb, emb_w, n_rows, E, topn, total_sq, max_norm, lr);
#else
    rows_sgd_kernel<<<n_rows, kEmbThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(words), owner,
                                                                      g_emb, emb_w, n_rows, E, topn, total_sq, max_norm,
                                                                      lr);
#endif
    return gpt_launch_status();
}
