"""Tensor-level wrappers and autograd Functions over the C ABI (include/gpt_b200.h).

PyTorch is plumbing here: it owns device memory and the stream, and autograd orders the backward calls so that
the caller-owned ``loss.backward()`` of the reference's train loop (/root/reference/train.py:220-221) keeps working.
All arithmetic of the hot path happens in libgptb200.so.
"""
import torch

from . import _lib

POOL_TYPES = {'max': 0, 'avg': 1, 'sum': 2}
GEMM_MODES = {'fp32': 0, 'tf32': 1, 'tf32x3': 2, 'bf16': 3}

TREE_ERR_FATAL = 1 | 2 | 4 | 8 | 16 | 32
TREE_ERR_NAMES = {1: 'head out of range', 2: 'no root', 4: 'cycle in head', 8: 'empty subject span',
                  16: 'entities under different roots', 32: 'deprel id out of range',
                  64: 'kept edge with deprel 0 (asymmetric adjacency)'}


def _stream():
    return torch.cuda.current_stream().cuda_stream


class KernelTimer(object):
    """Optional per-entry-point CUDA-event timing (bench.py's kernel table / roofline); off by default."""

    def __init__(self):
        self.events = {}

    def add(self, name, start, stop):
        self.events.setdefault(name, []).append((start, stop))

    def summary(self):
        """{name: (calls, total_ms)}; synchronises."""
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.events.items()}


TIMER = None   # set to a KernelTimer to time every C-ABI call on the launching stream


def _call(name, *args):
    fn = getattr(_lib.lib(), name)
    if TIMER is None:
        _lib.check(fn(*args), name)
        return
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    rc = fn(*args)
    stop.record()
    TIMER.add(name, start, stop)
    _lib.check(rc, name)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _dev(t, dtype, name):
    if not t.is_cuda:
        raise _lib.GptError('%s must be a CUDA tensor: the gpt_b200 path has no CPU fallback' % name)
    if t.dtype != dtype:
        raise _lib.GptError('%s must be %s, got %s' % (name, dtype, t.dtype))
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the current device's current stream (_stream()): a tensor of another GPU would be touched
        # from the wrong device.  Callers select the model's device (torch.cuda.set_device / torch.cuda.device).
        raise _lib.GptError('%s lives on %s but the current CUDA device is %d' % (name, t.device,
                                                                                 torch.cuda.current_device()))
    return t if t.is_contiguous() else t.contiguous()


class TreeCSR(object):
    """Device-resident pruned-tree adjacency of one padded batch (output of K1)."""
    __slots__ = ('B', 'T', 'rowptr', 'col', 'val', 'flags', 'denom', 'lens', 'err')

    def __init__(self, B, T, device):
        self.B, self.T = B, T
        self.rowptr = torch.empty((B, T + 1), dtype=torch.int32, device=device)
        self.col = torch.empty((B, 3 * T), dtype=torch.int32, device=device)
        self.val = torch.empty((B, 3 * T), dtype=torch.uint8, device=device)
        self.flags = torch.empty((B, T), dtype=torch.uint8, device=device)
        self.denom = torch.empty((B, T), dtype=torch.float32, device=device)
        self.lens = torch.empty((B,), dtype=torch.int32, device=device)
        self.err = torch.empty((B,), dtype=torch.int32, device=device)

    def to_dense(self):
        """[B,T,T] float32 with the reference's relation-id values (tests / debugging; synchronises)."""
        B, T = self.B, self.T
        rowptr = self.rowptr.cpu().long()
        col = self.col.cpu().long()
        val = self.val.cpu().float()
        adj = torch.zeros((B, T, T), dtype=torch.float32)
        counts = rowptr[:, 1:] - rowptr[:, :-1]
        for b in range(B):
            nnz = int(rowptr[b, T])
            if nnz == 0:
                continue
            rows = torch.repeat_interleave(torch.arange(T), counts[b])
            adj[b, rows, col[b, :nnz]] = val[b, :nnz]
        return adj

    def pool_mask(self):
        """bool [B,T,1], True where the token is not in the pruned tree (the reference's `mask`, gcn.py:262)."""
        return (self.flags & 1).eq(0).unsqueeze(2)

    def check(self, symmetric=False):
        """Raise on sentences the reference cannot process (synchronises; call outside the hot loop).  With
        ``symmetric`` also on a kept edge whose deprel id is 0: the forward still matches the reference's (asymmetric)
        matrix, but the backward kernels walk the CSR as its own transpose and would be wrong for that sentence --
        call it with symmetric=True on training data (the loader never emits id 0, data/loader.py:158-160)."""
        bad = (self.err & (TREE_ERR_FATAL | (64 if symmetric else 0))).nonzero().flatten().tolist()
        if bad:
            code = int(self.err[bad[0]])
            why = ', '.join(n for bit, n in TREE_ERR_NAMES.items() if code & bit)
            raise _lib.GptError('malformed dependency tree in batch row %d: %s' % (bad[0], why))


def prune_csr(head, subj_pos, obj_pos, deprel, masks, prune_k, out=None):
    """K1: batched head_to_tree + tree_to_adj (model/tree.py:58-204, model/gcn.py:96-110) -> TreeCSR."""
    head = _dev(head, torch.int64, 'head')
    subj_pos = _dev(subj_pos, torch.int64, 'subj_pos')
    obj_pos = _dev(obj_pos, torch.int64, 'obj_pos')
    deprel = _dev(deprel, torch.int64, 'deprel')
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8) if masks.is_contiguous() else masks.contiguous().view(torch.uint8)
    masks = _dev(masks, torch.uint8, 'masks')
    B, T = head.shape
    csr = TreeCSR(B, T, head.device) if out is None else out
    _call('gpt_prune_csr', _ptr(head), _ptr(subj_pos), _ptr(obj_pos), _ptr(deprel), _ptr(masks), B, T,
          int(prune_k), _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.val), _ptr(csr.flags),
          _ptr(csr.denom), _ptr(csr.lens), _ptr(csr.err), _stream())
    return csr


def l2_prefetch(t):
    """Warm L2 with a tensor's bytes ahead of its first use."""
    _call('gpt_l2_prefetch', _ptr(t), t.numel() * t.element_size(), _stream())


# ---- K3: projection ------------------------------------------------------------------------------------------

def _tf32_ok(*dims):
    return all(d % 4 == 0 for d in dims)


def gemm_persist_config(cta_group=2, min_rows=65536):
    """Which projections take the persistent large-M kernel (csrc/gemm_persist.cuh): cta_group 0 = never, 1 = single-CTA
    tiles, 2 = CTA pairs; from min_rows rows up.  Process-wide."""
    _lib.check(_lib.lib().gpt_gemm_persist_config(int(cta_group), int(min_rows)), 'gpt_gemm_persist_config')


def _bf16_ok(*dims):
    return all(d % 8 == 0 for d in dims)


def weight_prep_buffer(weight, mode):
    N, K = weight.shape
    if mode == 'bf16' and _bf16_ok(N, K):
        return torch.empty((2, N * K), dtype=torch.bfloat16, device=weight.device)      # [bf16(w) | bf16(w^T)]
    if mode != 'tf32x3' or not _tf32_ok(N, K):
        return None
    return torch.empty((4, N * K), dtype=torch.float32, device=weight.device)


def weight_prep(weight, mode, out=None):
    """Per-step operand preparation: 3xTF32 -> [w_hi | w_lo | w^T_hi | w^T_lo] (fp32); bf16 -> [bf16(w) | bf16(w^T)];
    None for the modes / shapes that read the weight as it is."""
    N, K = weight.shape
    ws = weight_prep_buffer(weight, mode) if out is None else out
    if ws is None:
        return None
    _call('gpt_weight_prep_bf16' if ws.dtype == torch.bfloat16 else 'gpt_weight_prep_tf32x3', _ptr(weight), _ptr(ws), N,
          K, _stream())
    return ws


def weight_prep_all(weights, mode, outs):
    """weight_prep for every layer in one launch; layers whose shape has no tensor-core path (out is None) are skipped."""
    import ctypes
    todo = [(w, o) for w, o in zip(weights, outs) if o is not None]
    if mode not in ('tf32x3', 'bf16') or not todo:
        return
    n = len(todo)
    if n > 8 or mode == 'bf16':
        for w, o in todo:
            weight_prep(w, mode, out=o)
        return
    _call('gpt_weight_prep_tf32x3_batch', (ctypes.c_void_p * n)(*[_ptr(w) for w, _ in todo]),
          (ctypes.c_void_p * n)(*[_ptr(o) for _, o in todo]), (ctypes.c_int * n)(*[w.shape[0] for w, _ in todo]),
          (ctypes.c_int * n)(*[w.shape[1] for w, _ in todo]), n, _stream())


def linear_fwd(x2d, weight, mode='fp32', ws=None):
    """y = x W^T.  'tf32x3': tcgen05/TMEM GEMM fed by TMA, 3xTF32 (fp32-grade); 'tf32': one TF32 pass; 'bf16': operands
    rounded to bf16 (kind::f16), fp32 accumulation; 'fp32': FFMA.  The tensor-core modes fall back to FFMA when the shape
    cannot be described to TMA (row pitch % 16 B != 0)."""
    M, K = x2d.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x2d.device)
    if mode not in GEMM_MODES:
        raise _lib.GptError('unknown gemm mode %r' % mode)
    if mode in ('tf32x3', 'bf16') and ws is None:
        ws = weight_prep(weight, mode)
    if mode == 'tf32' and _tf32_ok(K):
        _call('gpt_linear_fwd_tf32', _ptr(x2d), _ptr(weight), _ptr(y), M, N, K, _stream())
    elif mode == 'tf32x3' and ws is not None:
        _call('gpt_linear_fwd_tf32x3', _ptr(x2d), _ptr(ws), _ptr(y), M, N, K, _stream())
    elif mode == 'bf16' and ws is not None:
        _call('gpt_linear_fwd_bf16', _ptr(x2d), _ptr(ws), _ptr(y), M, N, K, _stream())
    elif mode in ('fp32', 'tf32', 'tf32x3', 'bf16'):
        _call('gpt_linear_fwd_f32', _ptr(x2d), _ptr(weight), _ptr(y), M, N, K, _stream())
    else:
        raise _lib.GptError('gemm mode %r not built' % mode)
    return y


def linear_dgrad(dy, weight, mode='fp32', ws=None):
    M, N = dy.shape
    K = weight.shape[1]
    dx = torch.empty((M, K), dtype=torch.float32, device=dy.device)
    if mode in ('tf32x3', 'bf16') and ws is None:
        ws = weight_prep(weight, mode)
    if mode == 'tf32' and _tf32_ok(N, K):
        wt = torch.empty((K, N), dtype=torch.float32, device=dy.device)
        _call('gpt_linear_dgrad_tf32', _ptr(dy), _ptr(weight), _ptr(dx), _ptr(wt), M, N, K, _stream())
    elif mode == 'tf32x3' and ws is not None:
        _call('gpt_linear_dgrad_tf32x3', _ptr(dy), _ptr(ws), _ptr(dx), M, N, K, _stream())
    elif mode == 'bf16' and ws is not None:
        _call('gpt_linear_dgrad_bf16', _ptr(dy), _ptr(ws), _ptr(dx), M, N, K, _stream())
    else:
        _call('gpt_linear_dgrad_f32', _ptr(dy), _ptr(weight), _ptr(dx), M, N, K, _stream())
    return dx


import os as _os
WGRAD_TC_MIN_ROWS = int(_os.environ.get('GPT_WGRAD_TC_MIN_ROWS', 32768))   # below: FFMA kernel over the live rows


def wgrad_tc_ok(M, N, K):
    """Shapes csrc/wgrad_tcgen05.cu takes: TMA-describable, accumulator fits tensor memory, and enough work for its
    persistent CTAs -- a long reduction (the regular layers of large batches) or many 128-row slices of dW (the shared
    [D*H, in] projection of the relation-aware layers: 79 slices at D = 50, H = 200)."""
    return (M >= WGRAD_TC_MIN_ROWS or (M >= 1024 and N >= 2048)) and K % 4 == 0 and N % 4 == 0 and K <= 512


def linear_wgrad(dy, x2d, mode='fp32', out=None, accumulate=False, flags=None):
    """dw = dy^T x.  accumulate: add into ``out`` (which the caller keeps zeroed between steps) instead of overwriting;
    with ``flags`` (K1's per-row flags) only rows that carry a gradient are read."""
    M, N = dy.shape
    K = x2d.shape[1]
    if mode in ('tf32x3', 'bf16') and wgrad_tc_ok(M, N, K):
        # tensor cores (3xTF32), the row range / the slices of dW split over the SMs; the kernel adds into dw
        if out is None:
            dw = torch.zeros((N, K), dtype=torch.float32, device=dy.device)
        else:
            dw = out
            if not accumulate:
                dw.zero_()
        _call('gpt_linear_wgrad_tf32x3', _ptr(dy), _ptr(x2d), _ptr(flags), _ptr(dw), M, N, K, _stream())
        return dw
    dw = torch.empty((N, K), dtype=torch.float32, device=dy.device) if out is None else out
    if accumulate and flags is not None:
        _call('gpt_linear_wgrad_rows_f32', _ptr(dy), _ptr(x2d), _ptr(flags), _ptr(dw), M, N, K, _stream())
        return dw
    _call('gpt_linear_wgrad_f32_acc' if accumulate else 'gpt_linear_wgrad_f32', _ptr(dy), _ptr(x2d), _ptr(dw), M, N,
          K, _stream())
    return dw


# ---- K2: aggregation -----------------------------------------------------------------------------------------

def aggregate_pool_ok(B, T, H):
    """True when K2's last-layer forward can also produce the three max pools (csrc/aggregate.cu, POOL mode)."""
    return bool(_lib.lib().gpt_gcn_aggregate_fwd_pool_supported(int(B), int(T), int(H)))


def aggregate_bwd_pool(dpooled, argmax, act, csr, H, use_adj=True, dbias_out=None, live=None, compact_out=None):
    """K4 backward (max) + K2 backward of the last layer in one launch: d(pooled) [B,3H] -> dy [B*T, H].  With ``live``
    (LiveRows) and ``compact_out`` [B*T, H] the live rows of dy are also stored compactly (row inv[n] of compact_out)."""
    B, T = csr.B, csr.T
    dy = torch.empty((B * T, H), dtype=torch.float32, device=dpooled.device)
    _call('gpt_gcn_aggregate_bwd_pool_c', _ptr(dpooled), _ptr(argmax), _ptr(act), _ptr(csr.rowptr), _ptr(csr.col),
          _ptr(csr.denom), _ptr(dy), _ptr(dbias_out), _ptr(live.inv if live is not None else None),
          _ptr(compact_out if live is not None else None), B, T, H, int(bool(use_adj)), _stream())
    return dy


def aggregate_fwd_pool(y, csr, bias, use_adj=True, want_out=False):
    """K2 forward of the last layer fused with K4 (max): -> (pooled [B,3H], argmax [B,3H], act mask, out or None)."""
    B, T = csr.B, csr.T
    H = y.shape[-1]
    out = torch.empty((B, T, H), dtype=torch.float32, device=y.device) if want_out else None
    act = torch.empty((B * ((H + 31) // 32) * T,), dtype=torch.int32, device=y.device)
    pooled = torch.empty((B, 3 * H), dtype=torch.float32, device=y.device)
    argmax = torch.empty((B, 3 * H), dtype=torch.int32, device=y.device)
    _call('gpt_gcn_aggregate_fwd_pool', _ptr(y), _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.denom), _ptr(csr.flags),
          _ptr(bias), _ptr(out), _ptr(act), _ptr(pooled), _ptr(argmax), B, T, H, int(bool(use_adj)), _stream())
    return pooled, argmax, act, out


def aggregate_fwd(y, csr, bias, use_adj=True, drop_p=0.0, rng_state=None, subseq=0, drop_mask=None, force_vec=0,
                  want_act=False):
    """K2 forward; with want_act also returns the 1-bit-per-element activation mask the backward can read."""
    B, T = csr.B, csr.T
    H = y.shape[-1]
    out = torch.empty((B, T, H), dtype=torch.float32, device=y.device)
    act = torch.empty((B * ((H + 31) // 32) * T,), dtype=torch.int32, device=y.device) if want_act else None
    _call('gpt_gcn_aggregate_fwd', _ptr(y), _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.denom),
          _ptr(csr.flags), _ptr(bias), _ptr(out), _ptr(act), B, T, H, int(bool(use_adj)),
          float(drop_p), _ptr(rng_state), int(subseq), _ptr(drop_mask), int(force_vec), _stream())
    return (out, act) if want_act else out


def aggregate_bwd(gout, out, csr, use_adj=True, drop_p=0.0, drop_mask=None, want_dbias=True, force_vec=0, act=None,
                  dbias_out=None):
    """K2 backward; [out > 0] is taken from ``act`` (bit mask) when given, else from ``out``.  ``dbias_out``: add
    the bias gradient into this (caller-zeroed) buffer instead of a fresh one."""
    B, T = csr.B, csr.T
    H = gout.shape[-1]
    dy = torch.empty((B * T, H), dtype=torch.float32, device=gout.device)
    if dbias_out is not None:
        dbias = dbias_out
    else:
        dbias = torch.zeros((H,), dtype=torch.float32, device=gout.device) if want_dbias else None
    _call('gpt_gcn_aggregate_bwd', _ptr(gout), _ptr(out), _ptr(act), _ptr(csr.rowptr), _ptr(csr.col),
          _ptr(csr.denom), _ptr(dy), _ptr(dbias), B, T, H, int(bool(use_adj)), float(drop_p), _ptr(drop_mask),
          int(force_vec), _stream())
    return dy, dbias


class _GcnLayer(torch.autograd.Function):
    """One `regular` GCN layer (model/gcn.py:269-271, 390-393): projection GEMM + fused aggregation epilogue."""

    @staticmethod
    def forward(ctx, x, weight, bias, csr, use_adj, drop_p, rng_state, subseq, drop_mask, gemm_mode):
        x = _dev(x, torch.float32, 'x')
        weight = _dev(weight, torch.float32, 'weight')
        bias = _dev(bias, torch.float32, 'bias')
        if drop_mask is not None:
            drop_mask = _dev(drop_mask, torch.float32, 'drop_mask')
        B, T, K = x.shape
        ws = weight_prep(weight, gemm_mode)          # split / transposed weight, shared by forward and dgrad
        y = linear_fwd(x.view(B * T, K), weight, gemm_mode, ws)
        out, act = aggregate_fwd(y, csr, bias, use_adj, drop_p if drop_mask is None else 0.0, rng_state, subseq,
                                 drop_mask, want_act=True)
        ctx.csr, ctx.use_adj, ctx.gemm_mode = csr, use_adj, gemm_mode
        ctx.drop_p = drop_p if drop_mask is None else 0.0
        ctx.save_for_backward(x, weight, act, drop_mask, ws)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, weight, act, drop_mask, ws = ctx.saved_tensors
        B, T, K = x.shape
        gout = _dev(gout, torch.float32, 'grad_out')
        dy, dbias = aggregate_bwd(gout, None, ctx.csr, ctx.use_adj, ctx.drop_p, drop_mask,
                                  want_dbias=ctx.needs_input_grad[2], act=act)
        dx = linear_dgrad(dy, weight, ctx.gemm_mode, ws).view(B, T, K) if ctx.needs_input_grad[0] else None
        dw = linear_wgrad(dy, x.view(B * T, K), ctx.gemm_mode) if ctx.needs_input_grad[1] else None
        return dx, dw, dbias, None, None, None, None, None, None, None


def gcn_layer(x, weight, bias, csr, use_adj=True, drop_p=0.0, rng_state=None, subseq=0, drop_mask=None,
              gemm_mode='fp32'):
    return _GcnLayer.apply(x, weight, bias, csr, use_adj, drop_p, rng_state, subseq, drop_mask, gemm_mode)


# ---- K10: relation-aware layers (adj_type full_deprel / diagonal_deprel) ---------------------------------------------

class _Linear(torch.autograd.Function):
    """y = x W^T + b on the K3 GEMMs (the diagonal mode's preprocessor, model/gcn.py:153-155, 255-257)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gemm_mode):
        x = _dev(x, torch.float32, 'x')
        weight = _dev(weight, torch.float32, 'weight')
        ws = weight_prep(weight, gemm_mode)
        y = linear_fwd(x.view(-1, x.shape[-1]), weight, gemm_mode, ws)
        if bias is not None:
            y += bias
        ctx.gemm_mode = gemm_mode
        ctx.save_for_backward(x, weight, ws)
        return y.view(x.shape[:-1] + (weight.shape[0],))

    @staticmethod
    def backward(ctx, gy):
        x, weight, ws = ctx.saved_tensors
        gy2 = _dev(gy, torch.float32, 'grad_out').view(-1, weight.shape[0])
        dx = linear_dgrad(gy2, weight, ctx.gemm_mode, ws).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = linear_wgrad(gy2, x.view(-1, x.shape[-1]), ctx.gemm_mode) if ctx.needs_input_grad[1] else None
        db = None
        if ctx.needs_input_grad[2]:
            db = torch.zeros((weight.shape[0],), dtype=torch.float32, device=gy2.device)
            _call('gpt_colsum_acc', _ptr(gy2), gy2.shape[0], gy2.shape[1], _ptr(db), _stream())
        return dx, dw, db, None


def linear(x, weight, bias=None, gemm_mode='fp32'):
    return _Linear.apply(x, weight, bias, gemm_mode)


class RelationLayerConfig(object):
    """Per-layer switches of the relation-aware modes (all host scalars except the optional test masks)."""
    __slots__ = ('layer', 'deep', 'directed', 'self_loop', 'edge_keep', 'keep_edges', 'keep_tokens', 'drop_p',
                 'drop_mask', 'rng_state', 'gemm_mode', 'live')

    def __init__(self, layer, deep=False, directed=False, self_loop=True, edge_keep=1.0, keep_edges=None,
                 keep_tokens=None, drop_p=0.0, drop_mask=None, rng_state=None, gemm_mode='fp32', live=None):
        self.live = live                    # None or the batch's LiveRows: project / mix only the observable rows
        self.layer, self.deep, self.directed, self.self_loop = int(layer), bool(deep), bool(directed), bool(self_loop)
        self.edge_keep = float(edge_keep)
        self.keep_edges = keep_edges        # None or (uint8 [B,T,T] parent->child matrix, uint8 [B,T,T] child->parent)
        self.keep_tokens = keep_tokens      # None or (uint8 [B*T] forward, uint8 [B*T] reverse): 0 = relation forgotten
        self.drop_p, self.drop_mask, self.rng_state, self.gemm_mode = float(drop_p), drop_mask, rng_state, gemm_mode


def edge_keep_dense(rng_state, B, T, layer, direction, keep_prob):
    """The in-kernel edge-dropout decisions of (layer, direction) as a dense uint8 [B,T,T] (tests)."""
    out = torch.empty((B, T, T), dtype=torch.uint8, device=rng_state.device)
    _call('gpt_edge_keep_dense', _ptr(rng_state), B, T, int(layer), int(direction), float(keep_prob), _ptr(out),
          _stream())
    return out


def relation_keep_tokens(rng_state, n_rows, layer, keep_prop):
    """Relation forgetting draws (model/gcn.py:451-470): (keep_forward, keep_reverse) uint8 [n_rows]."""
    kf = torch.empty((n_rows,), dtype=torch.uint8, device=rng_state.device)
    kr = torch.empty((n_rows,), dtype=torch.uint8, device=rng_state.device)
    _call('gpt_relation_keep_tokens', _ptr(rng_state), int(n_rows), int(layer), float(keep_prop), _ptr(kf), _ptr(kr),
          _stream())
    return kf, kr


class LiveRows(object):
    """The observable rows of a batch (flags != 0: inside a pruned tree, or a subject / object token), compacted on the
    device by ``gpt_live_rows``: ``perm`` int32 [N] (ascending row ids, first ``count`` entries), ``inv`` int32 [N],
    ``live`` uint8 [N] (``i < count``), ``count`` int32 [1].  The count never comes to the host (graph capture).
    ``flags=None`` only allocates (on the current stream); ``run(flags)`` launches."""
    __slots__ = ('N', 'perm', 'inv', 'live', 'count')

    def __init__(self, flags=None, n_rows=None, device=None):
        if flags is not None:
            flags = _dev(flags, torch.uint8, 'flags')
            n_rows, device = flags.numel(), flags.device
        self.N = N = int(n_rows)
        self.perm = torch.empty((N,), dtype=torch.int32, device=device)
        self.inv = torch.empty((N,), dtype=torch.int32, device=device)
        self.live = torch.empty((N,), dtype=torch.uint8, device=device)
        self.count = torch.empty((1,), dtype=torch.int32, device=device)
        if flags is not None:
            self.run(flags)

    def run(self, flags):
        _call('gpt_live_rows', _ptr(flags), self.N, _ptr(self.perm), _ptr(self.inv), _ptr(self.live), _ptr(self.count),
              _stream())
        return self

    def gather(self, x2d, out=None):
        """[N,K] -> compact [N,K]: row i < count is x2d[perm[i]]; the rest is never read."""
        out = torch.empty_like(x2d) if out is None else out
        _call('gpt_gather_rows', _ptr(x2d), _ptr(self.perm), _ptr(self.count), self.N, x2d.shape[1], _ptr(out), _stream())
        return out

    def scatter(self, xc):
        """compact [N,K] -> [N,K] with zeros on the rows that are not observable."""
        out = torch.empty_like(xc)
        _call('gpt_scatter_rows', _ptr(xc), _ptr(self.inv), self.N, xc.shape[1], _ptr(out), _stream())
        return out


def wgrad_rows_tc_ok(N, K):
    """Shapes the tensor-core weight gradient over compacted rows takes (csrc/wgrad_tcgen05.cu)."""
    return K % 4 == 0 and N % 4 == 0 and K <= 512


def linear_wgrad_live_acc(dyc, xc, live, out):
    """out += dyc^T xc over the compact rows i < count (3xTF32 on the tensor cores)."""
    M, N = dyc.shape
    _call('gpt_linear_wgrad_tf32x3_rows', _ptr(dyc), _ptr(xc), _ptr(live.live), _ptr(out), M, N, xc.shape[1],
          _ptr(live.count), _stream())
    return out


def _linear_fwd_rows(xc, weight, mode, ws, live, bias=None):
    """-> (y, bias_added): over the live rows; the tensor-core path adds ``bias`` in its epilogue."""
    M, K = xc.shape
    N = weight.shape[0]
    if mode == 'tf32x3' and ws is not None and M < 65536:
        fuse = bias is not None and N % 4 == 0 and bias.data_ptr() % 16 == 0
        y = torch.empty((M, N), dtype=torch.float32, device=xc.device)
        _call('gpt_linear_fwd_tf32x3_rows', _ptr(xc), _ptr(ws), _ptr(bias if fuse else None), _ptr(y), M, N, K,
              _ptr(live.count), _stream())
        return y, fuse
    return linear_fwd(xc, weight, mode, ws), False          # every row: the ones beyond count are scratch


def _linear_dgrad_rows(dyc, weight, mode, ws, live):
    M, N = dyc.shape
    K = weight.shape[1]
    if mode == 'tf32x3' and ws is not None and M < 65536:
        dx = torch.empty((M, K), dtype=torch.float32, device=dyc.device)
        _call('gpt_linear_dgrad_tf32x3_rows', _ptr(dyc), _ptr(ws), _ptr(dx), M, N, K, _ptr(live.count), _stream())
        return dx
    return linear_dgrad(dyc, weight, mode, ws)


def _linear_wgrad_rows(dyc, xc, mode, live):
    """dw = dyc^T xc over the compact rows i < count (the rows beyond hold scratch: `live.live` masks them)."""
    M, N = dyc.shape
    K = xc.shape[1]
    dw = torch.zeros((N, K), dtype=torch.float32, device=dyc.device)
    if mode in ('tf32x3', 'bf16') and wgrad_tc_ok(M, N, K):
        _call('gpt_linear_wgrad_tf32x3_rows', _ptr(dyc), _ptr(xc), _ptr(live.live), _ptr(dw), M, N, K, _ptr(live.count),
              _stream())
    else:
        _call('gpt_linear_wgrad_rows_f32', _ptr(dyc), _ptr(xc), _ptr(live.live), _ptr(dw), M, N, K, _stream())
    return dw


def _agg3_fwd(F, R, S, csr, cfg):
    B, T = csr.B, csr.T
    H = F.shape[-1]
    out = torch.empty((B, T, H), dtype=torch.float32, device=F.device)
    kf, kr = cfg.keep_edges if cfg.keep_edges is not None else (None, None)
    _call('gpt_agg3_fwd', _ptr(F), _ptr(R), _ptr(S), _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.val), _ptr(csr.denom),
          _ptr(csr.flags), _ptr(kf), _ptr(kr), cfg.edge_keep, _ptr(cfg.rng_state), cfg.layer, int(cfg.directed),
          int(cfg.self_loop), cfg.drop_p if cfg.drop_mask is None else 0.0, _ptr(cfg.drop_mask), B, T, H, _ptr(out),
          _stream())
    return out


def _agg3_bwd(gout, out, csr, cfg):
    B, T = csr.B, csr.T
    H = gout.shape[-1]
    dF, dR, dS = (torch.empty((B * T, H), dtype=torch.float32, device=gout.device) for _ in range(3))
    kf, kr = cfg.keep_edges if cfg.keep_edges is not None else (None, None)
    _call('gpt_agg3_bwd', _ptr(gout), _ptr(out), _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.val), _ptr(csr.denom),
          _ptr(csr.flags), _ptr(kf), _ptr(kr), cfg.edge_keep, _ptr(cfg.rng_state), cfg.layer, int(cfg.directed),
          int(cfg.self_loop), cfg.drop_p if cfg.drop_mask is None else 0.0, _ptr(cfg.drop_mask), B, T, H, _ptr(dF),
          _ptr(dR), _ptr(dS), _stream())
    return dF, dR, dS


class _RelationLayerFull(torch.autograd.Function):
    """One full_deprel layer (model/gcn.py:296-386 + 390-393): ONE projection GEMM Z = x Wmat^T shared by the three
    directions, relation mix, CSR aggregation with the layer epilogue (csrc/deprel.cu)."""

    @staticmethod
    def forward(ctx, x, wmat, bias, emb, csr, deprel, cfg, ws):
        x = _dev(x, torch.float32, 'x')
        wmat = _dev(wmat, torch.float32, 'wmat')
        bias = _dev(bias, torch.float32, 'bias')
        emb = _dev(emb, torch.float32, 'deprel_emb.weight')
        deprel = _dev(deprel, torch.int64, 'deprel')
        B, T, K = x.shape
        D = emb.shape[1]
        H = wmat.shape[0] // D
        N = B * T
        if emb.shape[0] != 85 or wmat.shape != (D * H, K) or bias.numel() != D * H:
            raise _lib.GptError('full_deprel layer: inconsistent shapes')
        live = cfg.live
        # with `live`: xs / Z are COMPACT (row i < count belongs to token perm[i]); F, R, S and out stay per token
        xs = x.view(N, K) if live is None else live.gather(x.view(N, K))
        if live is None:
            Z, biased = linear_fwd(xs, wmat, cfg.gemm_mode, ws), False
        else:
            Z, biased = _linear_fwd_rows(xs, wmat, cfg.gemm_mode, ws, live, bias.reshape(-1))
        mix_bias = None if biased else bias       # `biased`: Z already holds x W^T + b, the mix must not add it again
        F, R, S = (torch.empty((N, H), dtype=torch.float32, device=x.device) for _ in range(3))
        kf, kr = cfg.keep_tokens if cfg.keep_tokens is not None else (None, None)
        _call('gpt_relmix_fwd_rows', _ptr(Z), _ptr(mix_bias), _ptr(emb), _ptr(deprel), _ptr(csr.flags), _ptr(kf), _ptr(kr),
              _ptr(None if live is None else live.perm), _ptr(None if live is None else live.count), N, D, H,
              int(cfg.deep), _ptr(F), _ptr(R), _ptr(S), _stream())
        out = _agg3_fwd(F, R, S, csr, cfg)
        ctx.csr, ctx.cfg, ctx.dims, ctx.biased = csr, cfg, (B, T, K, D, H), biased
        ctx.save_for_backward(xs, wmat, bias, emb, deprel, Z, out, ws)
        return out

    @staticmethod
    def backward(ctx, gout):
        xs, wmat, bias, emb, deprel, Z, out, ws = ctx.saved_tensors
        csr, cfg = ctx.csr, ctx.cfg
        live = cfg.live
        B, T, K, D, H = ctx.dims
        N = B * T
        gout = _dev(gout, torch.float32, 'grad_out')
        dF, dR, dS = _agg3_bwd(gout, out, csr, cfg)
        dZ = torch.empty((N, D * H), dtype=torch.float32, device=gout.device)     # compact, like Z, when `live`
        dE = torch.zeros_like(emb)
        kf, kr = cfg.keep_tokens if cfg.keep_tokens is not None else (None, None)
        cnt = None if live is None else live.count
        _call('gpt_relmix_bwd_rows', _ptr(Z), _ptr(None if ctx.biased else bias), _ptr(emb), _ptr(deprel), _ptr(csr.flags),
              _ptr(kf), _ptr(kr),
              _ptr(None if live is None else live.perm), _ptr(cnt), _ptr(dF), _ptr(dR), _ptr(dS), N, D, H, int(cfg.deep),
              _ptr(dZ), _ptr(dE), _stream())
        dbias = None
        if ctx.needs_input_grad[2]:
            dbias = torch.zeros((D * H,), dtype=torch.float32, device=gout.device)
            _call('gpt_colsum_acc_rows', _ptr(dZ), N, D * H, _ptr(cnt), _ptr(dbias), _stream())
        dx = dw = None
        if live is None:
            if ctx.needs_input_grad[0]:
                dx = linear_dgrad(dZ, wmat, cfg.gemm_mode, ws).view(B, T, K)
            if ctx.needs_input_grad[1]:
                dw = linear_wgrad(dZ, xs, cfg.gemm_mode)
        else:
            if ctx.needs_input_grad[0]:
                dx = live.scatter(_linear_dgrad_rows(dZ, wmat, cfg.gemm_mode, ws, live)).view(B, T, K)
            if ctx.needs_input_grad[1]:
                dw = _linear_wgrad_rows(dZ, xs, cfg.gemm_mode, live)
        return dx, dw, dbias, (dE if ctx.needs_input_grad[3] else None), None, None, None, None


def relation_layer_full(x, wmat, bias, emb, csr, deprel, cfg, ws=None):
    """x [B,T,K]; wmat [D*H, K] with wmat[d*H+h, k] = W.weight.reshape(D,K,H)[d,k,h] (model/gcn.py:301); bias [D*H]."""
    if ws is None:
        ws = weight_prep(wmat, cfg.gemm_mode)
    return _RelationLayerFull.apply(x, wmat, bias, emb, csr, deprel, cfg, ws)


class _RelationLayerDiag(torch.autograd.Function):
    """One diagonal_deprel layer (model/gcn.py:272-294 + 390-393): elementwise relation gates + CSR aggregation."""

    @staticmethod
    def forward(ctx, x, emb, csr, deprel, cfg):
        x = _dev(x, torch.float32, 'x')
        emb = _dev(emb, torch.float32, 'deprel_emb.weight')
        deprel = _dev(deprel, torch.int64, 'deprel')
        B, T, H = x.shape
        N = B * T
        if emb.shape != (85, H):
            raise _lib.GptError('diagonal_deprel layer: deprel_emb must be [85, hidden_dim]')
        F, R, S = (torch.empty((N, H), dtype=torch.float32, device=x.device) for _ in range(3))
        _call('gpt_diagmix_fwd', _ptr(x), _ptr(emb), _ptr(deprel), _ptr(csr.flags), N, H, _ptr(F), _ptr(R), _ptr(S),
              _stream())
        out = _agg3_fwd(F, R, S, csr, cfg)
        ctx.csr, ctx.cfg = csr, cfg
        ctx.save_for_backward(x, emb, deprel, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, emb, deprel, out = ctx.saved_tensors
        B, T, H = x.shape
        N = B * T
        gout = _dev(gout, torch.float32, 'grad_out')
        dF, dR, dS = _agg3_bwd(gout, out, ctx.csr, ctx.cfg)
        dx = torch.empty((B, T, H), dtype=torch.float32, device=gout.device)
        dE = torch.zeros_like(emb)
        _call('gpt_diagmix_bwd', _ptr(x), _ptr(emb), _ptr(deprel), _ptr(ctx.csr.flags), _ptr(dF), _ptr(dR), _ptr(dS), N,
              H, _ptr(dx), _ptr(dE), _stream())
        return dx, dE, None, None, None


def relation_layer_diag(x, emb, csr, deprel, cfg):
    return _RelationLayerDiag.apply(x, emb, csr, deprel, cfg)


# ---- K4: pooling -----------------------------------------------------------------------------------------------

class _Pool3(torch.autograd.Function):
    """cat[pool(h, not-in-tree), pool(h, not-subj), pool(h, not-obj)] (model/gcn.py:116-121, 473-483)."""

    @staticmethod
    def forward(ctx, h, csr, pool_type):
        h = _dev(h, torch.float32, 'h')
        B, T, H = h.shape
        out = torch.empty((B, 3 * H), dtype=torch.float32, device=h.device)
        argmax = torch.empty((B, 3 * H), dtype=torch.int32, device=h.device) if pool_type == 0 else None
        _call('gpt_pool3_fwd', _ptr(h), _ptr(csr.flags), B, T, H, pool_type, _ptr(out), _ptr(argmax), _stream())
        ctx.csr, ctx.pool_type, ctx.shape = csr, pool_type, (B, T, H)
        ctx.save_for_backward(argmax)
        return out

    @staticmethod
    def backward(ctx, gout):
        (argmax,) = ctx.saved_tensors
        B, T, H = ctx.shape
        gout = _dev(gout, torch.float32, 'grad_out')
        dh = torch.empty((B, T, H), dtype=torch.float32, device=gout.device)
        _call('gpt_pool3_bwd', _ptr(gout), _ptr(argmax), _ptr(ctx.csr.flags), B, T, H, ctx.pool_type, _ptr(dh),
              _stream())
        return dh, None, None


def pool3(h, csr, pool_type='max'):
    return _Pool3.apply(h, csr, POOL_TYPES[pool_type])


def pool3_fwd(h, csr, pool_type):
    """K4 without autograd (engine.FusedTrainStep): -> (pooled [B,3H], argmax or None)."""
    B, T, H = h.shape
    out = torch.empty((B, 3 * H), dtype=torch.float32, device=h.device)
    argmax = torch.empty((B, 3 * H), dtype=torch.int32, device=h.device) if pool_type == 0 else None
    _call('gpt_pool3_fwd', _ptr(h), _ptr(csr.flags), B, T, H, pool_type, _ptr(out), _ptr(argmax), _stream())
    return out, argmax


def pool3_bwd(gout, argmax, csr, pool_type, H):
    B, T = csr.B, csr.T
    dh = torch.empty((B, T, H), dtype=torch.float32, device=gout.device)
    _call('gpt_pool3_bwd', _ptr(gout), _ptr(argmax), _ptr(csr.flags), B, T, H, pool_type, _ptr(dh), _stream())
    return dh


# ---- K5: input embeddings ------------------------------------------------------------------------------------

class SparseEmbeddingState(object):
    """Workspace of the row-sparse word-embedding gradient path (engine.GraphedTrainStep): G is a dense [V, E]
    buffer that is all-zero between steps, owner[w] the first token of the current batch that uses word w."""

    def __init__(self, weight, topn):
        V, E = weight.shape
        self.G = torch.zeros_like(weight)
        self.owner = torch.full((V,), 0x7fffffff, dtype=torch.int32, device=weight.device)
        self.sq = torch.zeros((1,), dtype=torch.float32, device=weight.device)
        self.topn = int(min(max(topn, 0), V))
        self.words = None           # the batch's word ids (static buffer), set by the forward


class _EmbedConcat(torch.autograd.Function):
    """dropout(cat[emb(words), pos_emb(pos), ner_emb(ner)]) (model/gcn.py:235-247) in one kernel each way."""

    @staticmethod
    def forward(ctx, words, pos, ner, emb_w, pos_w, ner_w, drop_p, rng_state, subseq, flags, topn, sparse):
        words = _dev(words, torch.int64, 'words')
        emb_w = _dev(emb_w, torch.float32, 'emb.weight')
        n_rows = words.numel()
        V, E = emb_w.shape
        Dp = pos_w.shape[1] if pos_w is not None else 0
        Dn = ner_w.shape[1] if ner_w is not None else 0
        x = torch.empty(words.shape + (E + Dp + Dn,), dtype=torch.float32, device=words.device)
        _call('gpt_embed_fwd', _ptr(words), _ptr(pos if Dp else None), _ptr(ner if Dn else None), _ptr(emb_w),
              _ptr(pos_w), _ptr(ner_w), _ptr(x), n_rows, V, E, Dp, Dn, float(drop_p),
              _ptr(rng_state if drop_p > 0 else None), int(subseq), _stream())
        ctx.meta = (n_rows, V, E, Dp, Dn, float(drop_p), rng_state, int(subseq), int(topn), sparse)
        ctx.shapes = (emb_w.shape, None if pos_w is None else pos_w.shape, None if ner_w is None else ner_w.shape)
        ctx.save_for_backward(words, pos, ner, flags)
        if sparse is not None:
            sparse.words = words
        return x

    @staticmethod
    def backward(ctx, dx):
        words, pos, ner, flags = ctx.saved_tensors
        n_rows, V, E, Dp, Dn, drop_p, rng_state, subseq, topn, sparse = ctx.meta
        dx = _dev(dx, torch.float32, 'grad_out')
        dev = dx.device
        want_emb, want_pos, want_ner = ctx.needs_input_grad[3], ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        g_pos = torch.zeros(ctx.shapes[1], dtype=torch.float32, device=dev) if (want_pos and Dp) else None
        g_ner = torch.zeros(ctx.shapes[2], dtype=torch.float32, device=dev) if (want_ner and Dn) else None
        if sparse is not None:          # side effect into the engine's all-zero-between-steps buffer
            g_emb, owner, ret_emb = (sparse.G if want_emb else None), sparse.owner, None
        else:
            g_emb = torch.zeros(ctx.shapes[0], dtype=torch.float32, device=dev) if want_emb else None
            owner, ret_emb = None, g_emb
        embed_bwd(dx, flags, words, pos if Dp else None, ner if Dn else None, g_emb, g_pos, g_ner, owner, V, E, topn,
                  drop_p, rng_state, subseq)
        return None, None, None, ret_emb, g_pos, g_ner, None, None, None, None, None, None


def embed_concat(words, pos, ner, emb_w, pos_w, ner_w, drop_p=0.0, rng_state=None, subseq=0, flags=None, topn=None,
                 sparse=None):
    V = emb_w.shape[0]
    topn = V if topn is None else int(min(max(topn, 0), V))      # train.py's default topn is 1e10
    return _EmbedConcat.apply(words, pos, ner, emb_w, pos_w, ner_w, drop_p, rng_state, subseq, flags, topn, sparse)


def embed_rows_sqnorm(state):
    V, E = state.G.shape
    _call('gpt_embed_rows_sqnorm', _ptr(state.words), _ptr(state.owner), _ptr(state.G), state.words.numel(), E,
          state.topn, _ptr(state.sq), _stream())


def embed_rows_sgd(state, weight, total_sq, max_norm, lr):
    V, E = state.G.shape
    _call('gpt_embed_rows_sgd', _ptr(state.words), _ptr(state.owner), _ptr(state.G), _ptr(weight),
          state.words.numel(), E, state.topn, _ptr(total_sq), float(max_norm), float(lr), _stream())


def embed_fwd(words, pos, ner, emb_w, pos_w, ner_w, drop_p, rng_state, subseq):
    """K5 forward without autograd (engine.FusedTrainStep)."""
    V, E = emb_w.shape
    Dp = pos_w.shape[1] if pos_w is not None else 0
    Dn = ner_w.shape[1] if ner_w is not None else 0
    x = torch.empty(words.shape + (E + Dp + Dn,), dtype=torch.float32, device=words.device)
    _call('gpt_embed_fwd', _ptr(words), _ptr(pos if Dp else None), _ptr(ner if Dn else None), _ptr(emb_w),
          _ptr(pos_w), _ptr(ner_w), _ptr(x), words.numel(), V, E, Dp, Dn, float(drop_p),
          _ptr(rng_state if drop_p > 0 else None), int(subseq), _stream())
    return x


def embed_fwd_prep(words, pos, ner, emb_w, pos_w, ner_w, drop_p, rng_state, subseq, weights, outs):
    """embed_fwd + weight_prep_all(weights, 'tf32x3', outs) in one launch (the front of engine.FusedTrainStep's graph).
    Returns None when some layer has no 3xTF32 workspace (the caller makes the two calls)."""
    import ctypes
    n = len(weights)
    if n == 0 or n > 8 or any(o is None or o.dtype != torch.float32 for o in outs):
        return None
    V, E = emb_w.shape
    Dp = pos_w.shape[1] if pos_w is not None else 0
    Dn = ner_w.shape[1] if ner_w is not None else 0
    x = torch.empty(words.shape + (E + Dp + Dn,), dtype=torch.float32, device=words.device)
    _call('gpt_embed_fwd_prep', _ptr(words), _ptr(pos if Dp else None), _ptr(ner if Dn else None), _ptr(emb_w),
          _ptr(pos_w), _ptr(ner_w), _ptr(x), words.numel(), V, E, Dp, Dn, float(drop_p),
          _ptr(rng_state if drop_p > 0 else None), int(subseq),
          (ctypes.c_void_p * n)(*[_ptr(w) for w in weights]), (ctypes.c_void_p * n)(*[_ptr(o) for o in outs]),
          (ctypes.c_int * n)(*[w.shape[0] for w in weights]), (ctypes.c_int * n)(*[w.shape[1] for w in weights]), n,
          _stream())
    return x


EMBED_GROUPED_MIN_ROWS = int(_os.environ.get('GPT_EMBED_GROUPED_MIN_ROWS', 65536))


def embed_bwd(dx, flags, words, pos, ner, g_emb, g_pos, g_ner, owner, V, E, topn, drop_p, rng_state, subseq):
    """K5 backward: scatter-add into caller-zeroed tables (any of g_emb / g_pos / g_ner may be None).  Large batches group
    the token rows by word first and sum each word's rows once (no floating-point atomics on the word table)."""
    Dp = g_pos.shape[1] if g_pos is not None else 0
    Dn = g_ner.shape[1] if g_ner is not None else 0
    n_rows = words.numel()
    if g_emb is not None and n_rows >= EMBED_GROUPED_MIN_ROWS and E % 4 == 0 and (E + Dp + Dn) % 4 == 0 and E <= 512 \
            and Dp + Dn <= 192:
        ws = torch.empty(int(_lib.lib().gpt_embed_bwd_grouped_workspace(n_rows, V)), dtype=torch.int32, device=dx.device)
        _call('gpt_embed_bwd_grouped', _ptr(dx), _ptr(flags), _ptr(words), _ptr(pos if Dp else None),
              _ptr(ner if Dn else None), _ptr(g_emb), _ptr(g_pos), _ptr(g_ner), _ptr(owner), n_rows, V, E, Dp, Dn,
              int(topn), float(drop_p), _ptr(rng_state if drop_p > 0 else None), int(subseq), _ptr(ws), _stream())
        return
    _call('gpt_embed_bwd', _ptr(dx), _ptr(flags), _ptr(words), _ptr(pos if Dp else None), _ptr(ner if Dn else None),
          _ptr(g_emb), _ptr(g_pos), _ptr(g_ner), _ptr(owner), words.numel(), V, E, Dp, Dn, int(topn), float(drop_p),
          _ptr(rng_state if drop_p > 0 else None), int(subseq), _stream())


# ---- K6: classifier head -------------------------------------------------------------------------------------------

def _ptr_array(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class HeadBuffers(object):
    """Outputs / workspace of K6 for a batch of B sentences."""

    def __init__(self, B, H, C, n_mlp, device):
        f = dict(dtype=torch.float32, device=device)
        self.logits = torch.empty((B, C), **f)
        self.loss_rows = torch.empty((B,), **f)
        self.acts = torch.empty((B, n_mlp, H), **f)
        self.dacts = torch.empty((B, n_mlp, H), **f)
        self.dlogits = torch.empty((B, C), **f)
        self.dpooled = torch.empty((B, 3 * H), **f)
        self.loss = torch.empty((), **f)


def head_fwd_bwd(pooled, labels, weights, biases, wc, bc, pooling_l2, buf, train=True):
    """K6: out_mlp + classifier + CE / pooling_l2 loss (+ data gradients when train) into ``buf`` (HeadBuffers)."""
    B, K0 = pooled.shape
    H, C = K0 // 3, wc.shape[0]
    _call('gpt_head_fwd_bwd', _ptr(pooled), _ptr(labels), _ptr_array(weights), _ptr_array(biases), len(weights),
          _ptr(wc), _ptr(bc), B, H, C, float(pooling_l2), int(bool(train)), _ptr(buf.logits), _ptr(buf.loss_rows),
          _ptr(buf.acts), _ptr(buf.dacts), _ptr(buf.dlogits), _ptr(buf.dpooled), _stream())
    return buf


def head_wgrad(pooled, buf, dws, dbs, dwc, dbc):
    """K6 weight gradients (overwrites dws / dbs / dwc / dbc) and the scalar loss -> buf.loss."""
    B, K0 = pooled.shape
    H, C = K0 // 3, dwc.shape[0]
    _call('gpt_head_wgrad', _ptr(pooled), _ptr(buf.acts), _ptr(buf.dacts), _ptr(buf.dlogits), _ptr(buf.loss_rows),
          B, H, C, len(dws), _ptr_array(dws), _ptr_array(dbs), _ptr(dwc), _ptr(dbc), _ptr(buf.loss), _stream())
    return buf.loss


class _HeadLoss(torch.autograd.Function):
    """K6 under autograd: loss = mean CE(classifier(out_mlp(pooled)), labels) + pooling_l2 * mean_b |pooled[b, :H]|^2.
    The forward runs K6's two launches, which already produce d loss / d pooled and every weight / bias gradient of the
    head; the backward only scales them by the incoming gradient (one multi-tensor launch)."""

    @staticmethod
    def forward(ctx, pooled, labels, pooling_l2, wc, bc, *wb):
        pooled = _dev(pooled, torch.float32, 'pooled')
        labels = _dev(labels, torch.int64, 'labels')
        weights = [_dev(w.detach(), torch.float32, 'out_mlp weight') for w in wb[0::2]]
        biases = [_dev(b.detach(), torch.float32, 'out_mlp bias') for b in wb[1::2]]
        wc_, bc_ = _dev(wc.detach(), torch.float32, 'classifier.weight'), _dev(bc.detach(), torch.float32, 'classifier.bias')
        B, H, C = pooled.shape[0], pooled.shape[1] // 3, wc_.shape[0]
        buf = HeadBuffers(B, H, C, len(weights), pooled.device)
        head_fwd_bwd(pooled, labels, weights, biases, wc_, bc_, pooling_l2, buf, train=True)
        dws, dbs = [torch.empty_like(w) for w in weights], [torch.empty_like(b) for b in biases]
        dwc, dbc = torch.empty_like(wc_), torch.empty_like(bc_)
        loss = head_wgrad(pooled, buf, dws, dbs, dwc, dbc)
        ctx.n = len(weights)
        ctx.save_for_backward(buf.dpooled, dwc, dbc, *[t for pair in zip(dws, dbs) for t in pair])
        ctx.mark_non_differentiable(buf.logits)
        return loss, buf.logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        grads = torch._foreach_mul(list(ctx.saved_tensors), gloss)
        return (grads[0], None, None, grads[1], grads[2]) + tuple(grads[3:])


def head_loss(pooled, labels, pooling_l2, wc, bc, *wb):
    """(loss, logits) of the classifier head on pooled [B,3H]; wb = out_mlp's (weight, bias) pairs in order."""
    return _HeadLoss.apply(pooled, labels, pooling_l2, wc, bc, *wb)


def predict_tail(logits, labels, dest, result):
    """K11: softmax + argmax + mean CE + un-sort (model/trainer.py:118-123) -> packed ``result`` (uint8 device buffer of
    gpt_predict_result_bytes(B, C) bytes: probs f32 [B,C] | predictions i32 [B] | loss f32)."""
    B, C = logits.shape
    _call('gpt_predict_tail', _ptr(logits), _ptr(labels), _ptr(dest), B, C, _ptr(result), _stream())


def predict_result_bytes(B, C):
    return int(_lib.lib().gpt_predict_result_bytes(int(B), int(C)))


# ---- K7: clip + SGD ---------------------------------------------------------------------------------------------------

def update_partials(n, n_rows):
    return int(_lib.lib().gpt_update_partials(int(n), int(n_rows)))


def update_sqnorm(flat_grad, sparse, partials):
    """Partial sums of g^2 over the flat dense gradient and (sparse: SparseEmbeddingState or None) the live word rows."""
    n_rows = sparse.words.numel() if sparse is not None else 0
    _call('gpt_update_sqnorm', _ptr(flat_grad), flat_grad.numel(), _ptr(sparse.words if sparse else None),
          _ptr(sparse.owner if sparse else None), _ptr(sparse.G if sparse else None), n_rows,
          sparse.G.shape[1] if sparse else 1, sparse.topn if sparse else 0, _ptr(partials), _stream())


def update_apply(flat_param, flat_grad, sparse, emb_weight, partials, max_norm, lr, grad_scale=1.0, total_norm=None,
                 step_counter=None):
    n_rows = sparse.words.numel() if sparse is not None else 0
    _call('gpt_update_apply', _ptr(flat_param), _ptr(flat_grad), flat_grad.numel(),
          _ptr(sparse.words if sparse else None), _ptr(sparse.owner if sparse else None),
          _ptr(sparse.G if sparse else None), _ptr(emb_weight if sparse else None), n_rows,
          sparse.G.shape[1] if sparse else 1, sparse.topn if sparse else 0, _ptr(partials), float(max_norm), float(lr),
          float(grad_scale), _ptr(total_norm), _ptr(step_counter), _stream())


# ---- K8: data-parallel exchange over peer memory -----------------------------------------------------------------

class ExchangeRegion(object):
    """One rank's exchange region (csrc/dp.cu): cudaMalloc'ed by the library and shareable through its cudaIpc handle, or
    (``buffer``) a caller-owned device allocation -- e.g. torch symmetric memory, which also comes with an NVSwitch
    multicast mapping."""

    def __init__(self, world, cap_rows, E, V, n_flat, buffer=None):
        import ctypes
        self.shape = (int(world), int(cap_rows), int(E), int(V), int(n_flat))
        nbytes = self.bytes_for(*self.shape)
        self.buffer = buffer
        if buffer is not None:
            if buffer.numel() * buffer.element_size() < nbytes or buffer.data_ptr() % 256:
                raise _lib.GptError('exchange buffer too small or misaligned')
            self.ptr, self.handle = buffer.data_ptr(), None
        else:
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.check(_lib.lib().gpt_dp_alloc(nbytes, ctypes.byref(ptr), handle), 'gpt_dp_alloc')
            self.ptr, self.handle = ptr.value, handle.raw
        self.nbytes = nbytes
        self.n_partials = int(_lib.lib().gpt_dp_partials(*self.shape))
        _call('gpt_dp_region_init', self.ptr, *self.shape, _stream())
        torch.cuda.current_stream().synchronize()

    @staticmethod
    def bytes_for(world, cap_rows, E, V, n_flat):
        nbytes = _lib.lib().gpt_dp_region_bytes(int(world), int(cap_rows), int(E), int(V), int(n_flat))
        if nbytes <= 0:
            raise _lib.GptError('bad exchange layout %r' % ((world, cap_rows, E, V, n_flat),))
        return int(nbytes)

    def free(self):
        if self.ptr and self.buffer is None:
            _lib.lib().gpt_dp_free(self.ptr)
        self.ptr = self.buffer = None


def open_peer_region(handle):
    """Map a peer rank's region (its 64-byte cudaIpc handle) into this process -> device pointer."""
    import ctypes
    ptr = ctypes.c_void_p()
    _lib.check(_lib.lib().gpt_dp_open(ctypes.create_string_buffer(handle, 64), ctypes.byref(ptr)), 'gpt_dp_open')
    return ptr.value


def dp_push(region_ptrs, rank, shape, flat_grad, sparse, multicast=None):
    """multicast: the NVSwitch multicast address of the W regions (one multimem.st reaches every rank) or None (one store
    per peer)."""
    import ctypes
    arr = (ctypes.c_void_p * len(region_ptrs))(*region_ptrs)
    n_rows = sparse.words.numel() if sparse is not None else 0
    tail = (_ptr(flat_grad), _ptr(sparse.G if sparse else None), _ptr(sparse.owner if sparse else None),
            _ptr(sparse.words if sparse else None), n_rows, sparse.topn if sparse else 0, _stream())
    if multicast:
        _call('gpt_dp_push_multicast', arr, int(multicast), rank, *shape, *tail)
    else:
        _call('gpt_dp_push', arr, rank, *shape, *tail)


def dp_signal(region_ptrs, rank, shape):
    """Raise this rank's flags in every region (only needed when dp_reduce is called with signal=False)."""
    import ctypes
    arr = (ctypes.c_void_p * len(region_ptrs))(*region_ptrs)
    _call('gpt_dp_signal', arr, rank, *shape, _stream())


def dp_reduce(region_ptrs, rank, shape, flat_grad, partials, signal=True):
    import ctypes
    arr = (ctypes.c_void_p * len(region_ptrs))(*region_ptrs)
    _call('gpt_dp_reduce', arr, rank, 1 if signal else 0, *shape, _ptr(flat_grad), _ptr(partials), _stream())


def dp_apply(region_ptr, shape, flat_param, flat_grad, emb_weight, partials, max_norm, lr, total_norm=None,
             step_counter=None):
    _call('gpt_dp_apply', region_ptr, *shape, _ptr(flat_param), _ptr(flat_grad), _ptr(emb_weight), _ptr(partials),
          float(max_norm), float(lr), _ptr(total_norm), _ptr(step_counter), _stream())


# ---- backward chain with K2's prologue fused into its producers -------------------------------------------------

def drop_scale(p):
    """The scale K2 applies to kept elements: keep-probability is quantised to 16 bits."""
    th = min(int(p * 65536.0 + 0.5), 65535)
    return 65536.0 / (65536.0 - th) if th > 0 else 1.0


def pool3_bwd_masked(gout, argmax, csr, pool_type, H, act, p_drop):
    """K4 backward writing g = dh * dropscale * [out > 0] / denom (input of aggregate_bwd_pre)."""
    B, T = csr.B, csr.T
    g = torch.empty((B, T, H), dtype=torch.float32, device=gout.device)
    _call('gpt_pool3_bwd_masked', _ptr(gout), _ptr(argmax), _ptr(csr.flags), _ptr(act), _ptr(csr.denom),
          float(drop_scale(p_drop)), B, T, H, pool_type, _ptr(g), _stream())
    return g


def linear_dgrad_masked(dy, weight, ws, act_prev, csr, p_drop_prev):
    """K3 dgrad (3xTF32) writing g of the previous layer instead of dx; None when the shape is not TMA-describable."""
    M, N = dy.shape
    K = weight.shape[1]
    if ws is None or ws.dtype != torch.float32 or not _tf32_ok(N, K):
        return None
    g = torch.empty((M, K), dtype=torch.float32, device=dy.device)
    _call('gpt_linear_dgrad_tf32x3_masked', _ptr(dy), _ptr(ws), _ptr(g), _ptr(act_prev), _ptr(csr.denom),
          float(drop_scale(p_drop_prev)), csr.T, M, N, K, _stream())
    return g


def aggregate_bwd_pre(g, csr, use_adj=True, dbias_out=None, force_vec=0, live=None, compact_out=None):
    B, T = csr.B, csr.T
    H = g.shape[-1]
    dy = torch.empty((B * T, H), dtype=torch.float32, device=g.device)
    _call('gpt_gcn_aggregate_bwd_pre_c', _ptr(g), _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.denom), _ptr(dy),
          _ptr(dbias_out), _ptr(live.inv if live is not None else None),
          _ptr(compact_out if live is not None else None), B, T, H, int(bool(use_adj)), int(force_vec), _stream())
    return dy


# ---- K9: device-resident batch builder ------------------------------------------------------------------------------

def build_batch(arena, offsets, labels, sel, B, T, word_dropout, seed, stream_id, out, masks, rels):
    """One loader batch from the token arena (csrc/batch.cu).  arena / out: 7 entries (int32 / int64 device tensors,
    None for an absent NER field) in the order words, pos, ner, deprel, head, subj_pos, obj_pos."""
    import ctypes
    for t in [a for a in arena if a is not None] + [offsets, labels, sel, masks, rels] + [o for o in out if o is not None]:
        if not t.is_cuda:
            raise _lib.GptError('build_batch needs CUDA tensors: the gpt_b200 path has no CPU fallback')
    a = (ctypes.c_void_p * 7)(*[_ptr(t) for t in arena])
    o = (ctypes.c_void_p * 7)(*[_ptr(t) for t in out])
    _call('gpt_build_batch', a, _ptr(offsets), _ptr(labels), _ptr(sel), int(B), int(T), float(word_dropout),
          int(seed) & 0xffffffffffffffff, int(stream_id) & 0xffffffffffffffff, o, _ptr(masks), _ptr(rels), _stream())


def build_batch_raw(arena_arr, offsets_ptr, labels_ptr, sel_ptr, B, T, word_dropout, seed, stream_id, out_arr,
                    masks_ptr, rels_ptr):
    """build_batch with every pointer already resolved (the loader caches them: this runs once per training step)."""
    _call('gpt_build_batch', arena_arr, offsets_ptr, labels_ptr, sel_ptr, int(B), int(T), float(word_dropout),
          int(seed) & 0xffffffffffffffff, int(stream_id) & 0xffffffffffffffff, out_arr, masks_ptr, rels_ptr, _stream())
