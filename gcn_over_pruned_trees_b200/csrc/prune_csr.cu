// K1 -- batched path-centric pruning of dependency trees, emitted as a per-sentence CSR adjacency.
//
// Replaces, for a whole padded batch and without leaving the device,
//   head_to_tree   /root/reference/model/tree.py:58-165
//   tree_to_adj    /root/reference/model/tree.py:167-204  (directed=False, self_loop=True)
//   the host loop  /root/reference/model/gcn.py:96-110    (6 D2H syncs + one dense [B,T,T] H2D copy)
//   adj != 0, denom = rowsum + 1, mask = (rowsum + colsum) == 0   /root/reference/model/gcn.py:260-262
//
// One warp owns one sentence; all per-sentence state lives in that warp's slice of shared memory, every
// synchronisation is a __syncwarp().  Integer work only; output is bit-exact (tests densify it and compare with
// the reference's float adjacency, values 1..84 included).
//
// Output layout (static shapes, CUDA-graph friendly): sentence b owns col/val[b*cap .. b*cap+cap), cap = 3*T
// (a tree row set has at most 3*n_kept - 2 entries); rowptr[b, 0..T] are offsets into that segment; columns are
// ascending inside a row, so the CSR is canonical and the aggregation's summation order is deterministic.
#include "gpt_common.cuh"
#include <cstdlib>

namespace {

constexpr int kWarpsPerCta = 4;

constexpr unsigned M_DEPREL = 0xffu;
constexpr unsigned M_SUBJ = 1u << 8;
constexpr unsigned M_OBJ = 1u << 9;
constexpr unsigned M_ENT = 1u << 10;   // entity token inside the sentence (t < len)
constexpr unsigned M_PATH = 1u << 11;  // on the dependency path (tree.py:126-127)
constexpr unsigned M_EDGE = 1u << 12;  // kept, not the pruned root: hangs under its head (tree.py:158-160)
constexpr unsigned M_INTREE = 1u << 13;

constexpr int E_HEAD_RANGE = 1;    // head outside [0, len] or self-referential
constexpr int E_NO_ROOT = 2;       // tree.py:164 assert
constexpr int E_CYCLE = 4;         // tree.py:91-94 would never return
constexpr int E_EMPTY_SUBJ = 8;    // tree.py:109/113 raise on cas=None
constexpr int E_DISJOINT = 16;     // tree.py:121-127 UnboundLocalError: entities under different roots
constexpr int E_DEPREL_RANGE = 32; // relation id does not fit the uint8 value array
constexpr int E_DEPREL_PAD = 64;   // warning only: a kept edge has deprel 0 -> forward entry is 0
constexpr int E_FATAL = E_HEAD_RANGE | E_NO_ROOT | E_CYCLE | E_EMPTY_SUBJ | E_DISJOINT | E_DEPREL_RANGE;

__global__ void __launch_bounds__(kWarpsPerCta * 32)
prune_csr_kernel(const long long* __restrict__ head, const long long* __restrict__ subj_pos,
                 const long long* __restrict__ obj_pos, const long long* __restrict__ deprel,
                 const unsigned char* __restrict__ pad, int B, int T, int prune_k, int cap,
                 int* __restrict__ rowptr, int* __restrict__ col, unsigned char* __restrict__ val,
                 unsigned char* __restrict__ flags, float* __restrict__ denom, int* __restrict__ lens,
                 int* __restrict__ err) {
    extern __shared__ int smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kWarpsPerCta + warp;
    if (b >= B) return;  // warp-uniform; no block-wide barrier is used below

    const int stride = T + 1;
    int* par = smem + (size_t)warp * 6 * stride;  // parent index, -1 at a root
    int* cnt = par + stride;                      // entity-subtree counts, later child counts / write cursors
    int* aux = cnt + stride;                      // root pointers, later common-child marks, later kept flags
    int* cptr = aux + stride;                     // children CSR offsets [T+1]
    int* clist = cptr + stride;                   // children, ascending per parent; before that: has-child marks
    unsigned* meta = reinterpret_cast<unsigned*>(clist + stride);

    const size_t row0 = (size_t)b * T;
    const long long* hb = head + row0;
    const long long* sb = subj_pos + row0;
    const long long* ob = obj_pos + row0;
    const long long* db = deprel + row0;
    const unsigned char* pb = pad + row0;

    // ---- lengths: len = #(masks == 0)   (gcn.py:96) -------------------------------------------------------
    int len = 0;
    for (int t = lane; t < T; t += 32) len += (pb[t] == 0);
    len = warp_sum_i(len);

    // ---- load + validate ----------------------------------------------------------------------------------
    int e = 0, n_ent = 0, n_subj = 0, last_root = -1;
    for (int t = lane; t < T; t += 32) {
        unsigned m = 0;
        const bool is_s = (sb[t] == 0), is_o = (ob[t] == 0);
        if (is_s) m |= M_SUBJ;
        if (is_o) m |= M_OBJ;
        int p = -1;
        if (t < len) {
            long long h = hb[t];
            if (h < 0 || h > len || h == t + 1) { e |= E_HEAD_RANGE; h = 0; }
            p = (int)h - 1;
            long long d = db[t];
            if (d < 0 || d > 255 - 42) { e |= E_DEPREL_RANGE; d = 0; }
            m |= (unsigned)d;
            if (is_s || is_o) { m |= M_ENT; ++n_ent; }
            if (is_s) ++n_subj;
            if (p < 0) last_root = t;  // tree.py:76-77: later roots overwrite
        }
        par[t] = p;
        meta[t] = m;
        cnt[t] = 0;
        aux[t] = (p < 0) ? t : p;
        clist[t] = 0;
    }
    last_root = warp_max_i(last_root);
    if (last_root < 0) e |= E_NO_ROOT;
    __syncwarp();

    // ---- root of every token by pointer jumping; anything that does not land on a root sits on a cycle -------
    {
        const int iters = 33 - __clz(max(len, 1));
        for (int it = 0; it < iters; ++it) {
            for (int t = lane; t < len; t += 32) aux[t] = aux[aux[t]];
            __syncwarp();
        }
        for (int t = lane; t < len; t += 32)
            if (par[aux[t]] >= 0) e |= E_CYCLE;
    }
    n_ent = warp_sum_i(n_ent);
    n_subj = warp_sum_i(n_subj);
    if (prune_k >= 0 && n_subj == 0) e |= E_EMPTY_SUBJ;
    e = warp_or_i(e);

    int root = last_root;
    if (!(e & E_FATAL)) {
        if (prune_k < 0) {
            // tree.py:67-79 + the BFS of tree.py:175-200: only the last root's component is ever visited
            for (int t = lane; t < len; t += 32) aux[t] = (aux[t] == last_root);
        } else {
            // cnt[j] = number of entity tokens in j's subtree (atomic walk up from every entity token)
            for (int t = lane; t < len; t += 32) {
                if (meta[t] & M_ENT)
                    for (int j = t; j >= 0; j = par[j]) atomicAdd(&cnt[j], 1);
            }
            __syncwarp();
            // common ancestors: cnt == n_ent (tree.py:85-109); the LCA is the one without a common child (:112-124)
            for (int t = lane; t < len; t += 32) aux[t] = 0;
            __syncwarp();
            for (int t = lane; t < len; t += 32)
                if (cnt[t] == n_ent && par[t] >= 0) aux[par[t]] = 1;
            __syncwarp();
            int lca = -1;
            for (int t = lane; t < len; t += 32)
                if (cnt[t] == n_ent && !aux[t]) lca = max(lca, t);
            lca = warp_max_i(lca);
            if (lca < 0) e |= E_DISJOINT;
            root = lca;
            __syncwarp();
            // path = (entity ancestor chains) - common + {lca}   (tree.py:126-127)
            for (int t = lane; t < len; t += 32) {
                const int c = cnt[t];
                if ((c >= 1 && c < n_ent) || t == lca) meta[t] |= M_PATH;
            }
            __syncwarp();
            // dist <= k  <=>  a path node is met within k steps up (tree.py:130-147)
            for (int t = lane; t < len; t += 32) {
                int keep = (meta[t] & M_PATH) ? 1 : 0;
                int j = t;
                for (int d = 0; d < prune_k && !keep; ++d) {
                    j = par[j];
                    if (j < 0) break;
                    if (meta[j] & M_PATH) keep = 1;
                }
                aux[t] = keep;
            }
        }
    }
    __syncwarp();

    const bool fatal = (e & E_FATAL) != 0;
    // ---- edges: every kept token except the pruned root hangs under its head (tree.py:149-160) -------------
    if (!fatal) {
        for (int t = lane; t < len; t += 32) cnt[t] = 0;
        __syncwarp();
        for (int t = lane; t < len; t += 32) {
            const int p = par[t];
            if (aux[t] && t != root && p >= 0) {
                meta[t] |= M_EDGE;
                clist[p] = 1;  // has a kept child -> gets a self loop (tree.py:190-192)
                if ((meta[t] & M_DEPREL) != 0) atomicAdd(&cnt[p], 1);
                else e |= E_DEPREL_PAD;  // A[p,t] = 0: entry absent, reverse entry (42) still present
            }
        }
        __syncwarp();
    }

    // ---- row sizes, exclusive scans (rows and children lists), denom, flags --------------------------------
    int* rp = rowptr + (size_t)b * (T + 1);
    int row_base = 0, child_base = 0;
    for (int c0 = 0; c0 < T; c0 += 32) {
        const int t = c0 + lane;
        int nn = 0, nc = 0;
        unsigned m = 0;
        if (t < T) {
            m = meta[t];
            if (!fatal && t < len) {
                const int edge = (m & M_EDGE) ? 1 : 0;
                const int in_tree = (edge || clist[t]) ? 1 : 0;
                nc = cnt[t];
                nn = nc + edge + in_tree;
                if (in_tree) m |= M_INTREE;
            }
        }
        const int incl_r = warp_incl_scan_i(nn, lane);
        const int incl_c = warp_incl_scan_i(nc, lane);
        if (t < T) {
            meta[t] = m;
            rp[t] = row_base + incl_r - nn;
            cptr[t] = child_base + incl_c - nc;
            denom[row0 + t] = (float)(nn + 1);  // rowsum(adj != 0) + 1   (gcn.py:261)
            flags[row0 + t] = (unsigned char)(((m & M_INTREE) ? GPT_FLAG_INTREE : 0u) |
                                              ((m & M_SUBJ) ? GPT_FLAG_SUBJ : 0u) | ((m & M_OBJ) ? GPT_FLAG_OBJ : 0u));
        }
        row_base += __shfl_sync(GPT_FULL_MASK, incl_r, 31);
        child_base += __shfl_sync(GPT_FULL_MASK, incl_c, 31);
    }
    if (lane == 0) {
        rp[T] = row_base;
        cptr[T] = child_base;
        lens[b] = len;
    }
    e = warp_or_i(e);
    if (lane == 0) err[b] = e;
    if (fatal) return;
    __syncwarp();

    // ---- children lists in ascending order: chunks of 32 tokens are visited in order, ranks inside a chunk
    //      come from match_any, so no sort and no order-dependent atomics -------------------------------------
    for (int t = lane; t < len; t += 32) cnt[t] = cptr[t];
    __syncwarp();
    for (int c0 = 0; c0 < len; c0 += 32) {
        const int t = c0 + lane;
        const bool has = (t < len) && (meta[t] & M_EDGE) && (meta[t] & M_DEPREL);
        const int p = has ? par[t] : -1 - lane;
        const unsigned grp = __match_any_sync(GPT_FULL_MASK, p);
        const int rank = __popc(grp & ((1u << lane) - 1u));
        int base = 0;
        if (has) {
            base = cnt[p];
            clist[base + rank] = t;
        }
        __syncwarp();
        if (has && rank == 0) cnt[p] = base + __popc(grp);
        __syncwarp();
    }

    // ---- emit rows: merge {sorted children, self, parent} ----------------------------------------------------
    int* cb = col + (size_t)b * cap;
    unsigned char* vb = val + (size_t)b * cap;
    for (int r = lane; r < len; r += 32) {
        const unsigned m = meta[r];
        if (!(m & M_INTREE)) continue;
        int w = rp[r];
        const int p = (m & M_EDGE) ? par[r] : -1;
        const unsigned char pv = (unsigned char)((m & M_DEPREL) + 42u);  // tree.py:186-188
        bool par_done = (p < 0), self_done = false;
        const int c_end = cptr[r + 1];
        for (int i = cptr[r]; i <= c_end; ++i) {
            const int c = (i < c_end) ? clist[i] : 0x7fffffff;
            // pending {parent, self} entries smaller than the next child column, in ascending order
            if (!par_done && p < r && p < c) { cb[w] = p; vb[w] = pv; ++w; par_done = true; }
            if (!self_done && r < c) { cb[w] = r; vb[w] = 84; ++w; self_done = true; }  // tree.py:190-192
            if (!par_done && self_done && p < c) { cb[w] = p; vb[w] = pv; ++w; par_done = true; }
            if (i < c_end) { cb[w] = c; vb[w] = (unsigned char)(meta[c] & M_DEPREL); ++w; }  // tree.py:184
        }
    }
}


// ---- the same, one CTA per sentence -----------------------------------------------------------------------------------
// Long sentences (T >= 256; BASELINE.json configs[4] has 512 tokens in every sentence): with one warp per sentence every
// phase is a serial walk over T/32 chunks and an SM holds 16 such warps (6 arrays of T+1 ints each) -- 13 % of the HBM
// peak at T = 512.  Here kBlockThreads threads share one sentence: a phase costs T / kBlockThreads chunk-steps, an SM
// holds 16 CTAs = 64 warps, and the only inherently ordered pass of the warp form (children lists built chunk by chunk
// with match_any) is replaced by an unordered fill + a sort of every row's own few children.  Same output, bit for bit.
constexpr int kBlockThreads = 128;
constexpr int kBlockWarps = kBlockThreads / 32;

enum { RED_SUM = 0, RED_MAX = 1, RED_OR = 2 };
template <int OP>
__device__ __forceinline__ int block_reduce_i(int v, int* scratch) {
    v = OP == RED_SUM ? warp_sum_i(v) : (OP == RED_MAX ? warp_max_i(v) : warp_or_i(v));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = scratch[0];
#pragma unroll
    for (int w = 1; w < kBlockWarps; ++w) {
        const int o = scratch[w];
        r = OP == RED_SUM ? r + o : (OP == RED_MAX ? max(r, o) : (r | o));
    }
    __syncthreads();    // scratch may be reused
    return r;
}

__global__ void __launch_bounds__(kBlockThreads)
prune_csr_block_kernel(const long long* __restrict__ head, const long long* __restrict__ subj_pos,
                       const long long* __restrict__ obj_pos, const long long* __restrict__ deprel,
                       const unsigned char* __restrict__ pad, int B, int T, int prune_k, int cap,
                       int* __restrict__ rowptr, int* __restrict__ col, unsigned char* __restrict__ val,
                       unsigned char* __restrict__ flags, float* __restrict__ denom, int* __restrict__ lens,
                       int* __restrict__ err) {
    extern __shared__ int smem[];
    __shared__ int scratch[kBlockWarps];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;

    const int stride = T + 1, n_chunks = (T + 31) / 32;
    int* par = smem;                 // parent index, -1 at a root
    int* cnt = par + stride;         // entity-subtree counts, later child counts / write cursors
    int* aux = cnt + stride;         // root pointers, later common-child marks, kept flags, finally CSR row starts
    int* cptr = aux + stride;        // children CSR offsets [T+1]
    int* clist = cptr + stride;      // children lists; before that: has-child marks
    unsigned* meta = reinterpret_cast<unsigned*>(clist + stride);
    int* ctot_r = reinterpret_cast<int*>(meta + stride);     // per 32-token chunk: row entries, then their exclusive scan
    int* ctot_c = ctot_r + n_chunks + 1;                     // per chunk: children, then their exclusive scan

    const size_t row0 = (size_t)b * T;
    const long long* hb = head + row0;
    const long long* sb = subj_pos + row0;
    const long long* ob = obj_pos + row0;
    const long long* db = deprel + row0;
    const unsigned char* pb = pad + row0;

    // ---- lengths (gcn.py:96) ----------------------------------------------------------------------------------
    int len = 0;
    for (int t = tid; t < T; t += kBlockThreads) len += (pb[t] == 0);
    len = block_reduce_i<RED_SUM>(len, scratch);

    // ---- load + validate ----------------------------------------------------------------------------------
    int e = 0, n_ent = 0, n_subj = 0, last_root = -1;
    for (int t = tid; t < T; t += kBlockThreads) {
        unsigned m = 0;
        const bool is_s = (sb[t] == 0), is_o = (ob[t] == 0);
        if (is_s) m |= M_SUBJ;
        if (is_o) m |= M_OBJ;
        int p = -1;
        if (t < len) {
            long long h = hb[t];
            if (h < 0 || h > len || h == t + 1) { e |= E_HEAD_RANGE; h = 0; }
            p = (int)h - 1;
            long long d = db[t];
            if (d < 0 || d > 255 - 42) { e |= E_DEPREL_RANGE; d = 0; }
            m |= (unsigned)d;
            if (is_s || is_o) { m |= M_ENT; ++n_ent; }
            if (is_s) ++n_subj;
            if (p < 0) last_root = max(last_root, t);  // tree.py:76-77: later roots overwrite
        }
        par[t] = p;
        meta[t] = m;
        cnt[t] = 0;
        aux[t] = (p < 0) ? t : p;
        clist[t] = 0;
    }
    last_root = block_reduce_i<RED_MAX>(last_root, scratch);   // (its barriers also publish the arrays)
    if (last_root < 0) e |= E_NO_ROOT;

    // ---- root of every token by pointer jumping (pointers only ever move towards the root, so reading a neighbour's
    //      pointer before or after its own update of the same round is equally good) --------------------------------
    {
        const int iters = 33 - __clz(max(len, 1));
        for (int it = 0; it < iters; ++it) {
            for (int t = tid; t < len; t += kBlockThreads) aux[t] = aux[aux[t]];
            __syncthreads();
        }
        for (int t = tid; t < len; t += kBlockThreads)
            if (par[aux[t]] >= 0) e |= E_CYCLE;
    }
    n_ent = block_reduce_i<RED_SUM>(n_ent, scratch);
    n_subj = block_reduce_i<RED_SUM>(n_subj, scratch);
    if (prune_k >= 0 && n_subj == 0) e |= E_EMPTY_SUBJ;
    e = block_reduce_i<RED_OR>(e, scratch);

    int root = last_root;
    if (!(e & E_FATAL)) {       // CTA-uniform
        if (prune_k < 0) {
            for (int t = tid; t < len; t += kBlockThreads) aux[t] = (aux[t] == last_root);
        } else {
            for (int t = tid; t < len; t += kBlockThreads) {
                if (meta[t] & M_ENT)
                    for (int j = t; j >= 0; j = par[j]) atomicAdd(&cnt[j], 1);
            }
            __syncthreads();
            for (int t = tid; t < len; t += kBlockThreads) aux[t] = 0;
            __syncthreads();
            for (int t = tid; t < len; t += kBlockThreads)
                if (cnt[t] == n_ent && par[t] >= 0) aux[par[t]] = 1;
            __syncthreads();
            int lca = -1;
            for (int t = tid; t < len; t += kBlockThreads)
                if (cnt[t] == n_ent && !aux[t]) lca = max(lca, t);
            lca = block_reduce_i<RED_MAX>(lca, scratch);
            if (lca < 0) e |= E_DISJOINT;
            root = lca;
            for (int t = tid; t < len; t += kBlockThreads) {
                const int c = cnt[t];
                if ((c >= 1 && c < n_ent) || t == lca) meta[t] |= M_PATH;
            }
            __syncthreads();
            for (int t = tid; t < len; t += kBlockThreads) {
                int keep = (meta[t] & M_PATH) ? 1 : 0;
                int j = t;
                for (int d = 0; d < prune_k && !keep; ++d) {
                    j = par[j];
                    if (j < 0) break;
                    if (meta[j] & M_PATH) keep = 1;
                }
                aux[t] = keep;
            }
        }
    }
    __syncthreads();

    const bool fatal = (e & E_FATAL) != 0;
    // ---- edges: every kept token except the pruned root hangs under its head (tree.py:149-160) -------------
    if (!fatal) {
        for (int t = tid; t < len; t += kBlockThreads) cnt[t] = 0;
        __syncthreads();
        for (int t = tid; t < len; t += kBlockThreads) {
            const int p = par[t];
            if (aux[t] && t != root && p >= 0) {
                meta[t] |= M_EDGE;
                clist[p] = 1;  // has a kept child -> gets a self loop (tree.py:190-192)
                if ((meta[t] & M_DEPREL) != 0) atomicAdd(&cnt[p], 1);
                else e |= E_DEPREL_PAD;  // A[p,t] = 0: entry absent, reverse entry (42) still present
            }
        }
        __syncthreads();
    }

    // ---- row sizes, denom, flags; exclusive scans in two levels: inside a chunk of 32 tokens, then over the chunks ---
    int* rp = rowptr + (size_t)b * (T + 1);
    for (int ch = warp; ch < n_chunks; ch += kBlockWarps) {
        const int t = ch * 32 + lane;
        int nn = 0, nc = 0;
        unsigned m = 0;
        if (t < T) {
            m = meta[t];
            if (!fatal && t < len) {
                const int edge = (m & M_EDGE) ? 1 : 0;
                const int in_tree = (edge || clist[t]) ? 1 : 0;
                nc = cnt[t];
                nn = nc + edge + in_tree;
                if (in_tree) m |= M_INTREE;
            }
        }
        const int incl_r = warp_incl_scan_i(nn, lane);
        const int incl_c = warp_incl_scan_i(nc, lane);
        if (t < T) {
            meta[t] = m;
            aux[t] = incl_r - nn;       // + the chunk's base below
            cptr[t] = incl_c - nc;
            denom[row0 + t] = (float)(nn + 1);  // rowsum(adj != 0) + 1   (gcn.py:261)
            flags[row0 + t] = (unsigned char)(((m & M_INTREE) ? GPT_FLAG_INTREE : 0u) |
                                              ((m & M_SUBJ) ? GPT_FLAG_SUBJ : 0u) | ((m & M_OBJ) ? GPT_FLAG_OBJ : 0u));
        }
        if (lane == 31) { ctot_r[ch] = incl_r; ctot_c[ch] = incl_c; }
    }
    __syncthreads();
    if (warp == 0) {
        int base_r = 0, base_c = 0;
        for (int c0 = 0; c0 < n_chunks; c0 += 32) {
            const int ch = c0 + lane;
            const int vr = ch < n_chunks ? ctot_r[ch] : 0, vc = ch < n_chunks ? ctot_c[ch] : 0;
            const int ir = warp_incl_scan_i(vr, lane), ic = warp_incl_scan_i(vc, lane);
            if (ch < n_chunks) { ctot_r[ch] = base_r + ir - vr; ctot_c[ch] = base_c + ic - vc; }
            base_r += __shfl_sync(GPT_FULL_MASK, ir, 31);
            base_c += __shfl_sync(GPT_FULL_MASK, ic, 31);
        }
        if (lane == 0) {
            rp[T] = base_r;
            cptr[T] = base_c;
            lens[b] = len;
        }
    }
    __syncthreads();
    for (int t = tid; t < T; t += kBlockThreads) {
        const int rs = aux[t] + ctot_r[t >> 5];
        aux[t] = rs;
        rp[t] = rs;
        cptr[t] += ctot_c[t >> 5];
    }
    e = block_reduce_i<RED_OR>(e, scratch);     // (its barriers also publish aux / cptr)
    if (tid == 0) err[b] = e;
    if (fatal) return;

    // ---- children lists: filled in any order, then every row sorts its own (short) list ----------------------
    for (int t = tid; t < len; t += kBlockThreads) cnt[t] = cptr[t];
    __syncthreads();
    for (int t = tid; t < len; t += kBlockThreads) {
        const unsigned m = meta[t];
        if ((m & M_EDGE) && (m & M_DEPREL)) clist[atomicAdd(&cnt[par[t]], 1)] = t;
    }
    __syncthreads();

    // ---- emit rows: merge {sorted children, self, parent} ----------------------------------------------------
    int* cb = col + (size_t)b * cap;
    unsigned char* vb = val + (size_t)b * cap;
    for (int r = tid; r < len; r += kBlockThreads) {
        const unsigned m = meta[r];
        if (!(m & M_INTREE)) continue;
        const int c_beg = cptr[r], c_end = cptr[r + 1];
        for (int i = c_beg + 1; i < c_end; ++i) {       // insertion sort: this thread is the segment's only user
            const int v = clist[i];
            int j = i - 1;
            while (j >= c_beg && clist[j] > v) { clist[j + 1] = clist[j]; --j; }
            clist[j + 1] = v;
        }
        int w = aux[r];
        const int p = (m & M_EDGE) ? par[r] : -1;
        const unsigned char pv = (unsigned char)((m & M_DEPREL) + 42u);  // tree.py:186-188
        bool par_done = (p < 0), self_done = false;
        for (int i = c_beg; i <= c_end; ++i) {
            const int c = (i < c_end) ? clist[i] : 0x7fffffff;
            // pending {parent, self} entries smaller than the next child column, in ascending order
            if (!par_done && p < r && p < c) { cb[w] = p; vb[w] = pv; ++w; par_done = true; }
            if (!self_done && r < c) { cb[w] = r; vb[w] = 84; ++w; self_done = true; }  // tree.py:190-192
            if (!par_done && self_done && p < c) { cb[w] = p; vb[w] = pv; ++w; par_done = true; }
            if (i < c_end) { cb[w] = c; vb[w] = (unsigned char)(meta[c] & M_DEPREL); ++w; }  // tree.py:184
        }
    }
}

}  // namespace

extern "C" int gpt_prune_csr(const int64_t* head, const int64_t* subj_pos, const int64_t* obj_pos,
                             const int64_t* deprel, const uint8_t* pad_mask, int B, int T, int prune_k,
                             int32_t* rowptr, int32_t* col, uint8_t* val, uint8_t* flags, float* denom,
                             int32_t* lens, int32_t* err, void* stream) {
    GPT_CHECK_ARG(head && subj_pos && obj_pos && deprel && pad_mask && rowptr && col && val && flags && denom &&
                  lens && err);
    GPT_CHECK_ARG(B >= 0 && T >= 1);
    if (B == 0) return GPT_OK;
    int block_min_t = 256;                                      // from here on: one CTA per sentence
    if (const char* e = getenv("GPT_K1_BLOCK_MIN_T")) block_min_t = atoi(e);      // tuning / test knob
    if (T >= block_min_t) {
        const size_t smem_b = ((size_t)6 * (T + 1) + 2 * ((T + 31) / 32 + 1)) * sizeof(int);
        if (smem_b > 200 * 1024) return GPT_ERR_UNSUPPORTED;    // T <= 8500 per sentence
        if (int a = gpt_smem_opt_in(prune_csr_block_kernel, smem_b)) return a;
        // (a plain launch, like the warp form below: K1 has no griddepcontrol.wait, it must not start early)
#ifdef GPT_HOST_EMULATION   // tests/emu: g++ has no <<<>>>
        gpt_launch(prune_csr_block_kernel, dim3(B), dim3(kBlockThreads), smem_b, (cudaStream_t)stream,
#else
        prune_csr_block_kernel<<<B, kBlockThreads, smem_b, (cudaStream_t)stream>>>(
#endif
            reinterpret_cast<const long long*>(head), reinterpret_cast<const long long*>(subj_pos),
            reinterpret_cast<const long long*>(obj_pos), reinterpret_cast<const long long*>(deprel), pad_mask, B, T,
            prune_k, 3 * T, rowptr, col, val, flags, denom, lens, err);
        return gpt_launch_status();
    }
    const size_t smem = (size_t)kWarpsPerCta * 6 * (T + 1) * sizeof(int);
    if (smem > 200 * 1024) return GPT_ERR_UNSUPPORTED;  // T <= 2132 per sentence
    if (int a = gpt_smem_opt_in(prune_csr_kernel, smem)) return a;
    const int grid = (B + kWarpsPerCta - 1) / kWarpsPerCta;
#ifdef GPT_HOST_EMULATION   // tests/emu: g++ has no <<<>>>
    gpt_launch(prune_csr_kernel, dim3(grid), dim3(kWarpsPerCta * 32), smem, (cudaStream_t)stream,
#else
    prune_csr_kernel<<<grid, kWarpsPerCta * 32, smem, (cudaStream_t)stream>>>(
#endif
        reinterpret_cast<const long long*>(head), reinterpret_cast<const long long*>(subj_pos),
        reinterpret_cast<const long long*>(obj_pos), reinterpret_cast<const long long*>(deprel), pad_mask, B, T,
        prune_k, 3 * T, rowptr, col, val, flags, denom, lens, err);
    return gpt_launch_status();
}
