"""GPU tests of the fused step: K6 (classifier head), K7 (clip + SGD) and engine.FusedTrainStep, against plain PyTorch
fp32 restatements of the reference lines they replace (model/gcn.py:64-68,122, model/trainer.py:94-100,
train.py:224-227) and against the autograd path of this package.

Tolerances: logits / loss <= 1e-5 relative, gradients <= 1e-4 relative per tensor, parameters after N steps <= 5e-5.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gcn_over_pruned_trees_b200 import ops, synth
from gcn_over_pruned_trees_b200.engine import FusedTrainStep, GraphedTrainStep
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    a = a.detach().double().cpu().numpy()
    b = b.detach().double().cpu().numpy()
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------- K6 ------------------------------------------------

@pytest.mark.parametrize('B,H,C,L,l2', [(50, 200, 42, 2, 0.003), (7, 64, 10, 1, 0.0), (300, 200, 19, 3, 0.003),
                                        (1300, 32, 42, 2, 0.01), (1, 200, 42, 2, 0.003), (64, 512, 42, 2, 0.003),
                                        (33, 36, 5, 4, 0.5)])
def test_k6_head_matches_torch(B, H, C, L, l2):
    g = torch.Generator().manual_seed(B * 7 + H)
    pooled = (torch.randn(B, 3 * H, generator=g) * 1.5).to(DEV).requires_grad_()
    labels = torch.randint(0, C, (B,), generator=g).to(DEV)
    ws = [(torch.randn(H, 3 * H if l == 0 else H, generator=g) / np.sqrt(3 * H if l == 0 else H)).to(DEV).requires_grad_()
          for l in range(L)]
    bs = [(torch.randn(H, generator=g) * 0.1).to(DEV).requires_grad_() for _ in range(L)]
    wc = (torch.randn(C, H, generator=g) / np.sqrt(H)).to(DEV).requires_grad_()
    bc = (torch.randn(C, generator=g) * 0.1).to(DEV).requires_grad_()
    # reference: out_mlp -> classifier -> CE + pooling_l2 * mean_b sum_h h_out^2
    h = pooled
    for w, b in zip(ws, bs):
        h = F.relu(F.linear(h, w, b))
    logits = F.linear(h, wc, bc)
    loss = F.cross_entropy(logits, labels) + l2 * (pooled[:, :H] ** 2).sum(1).mean()
    loss.backward()

    buf = ops.HeadBuffers(B, H, C, L, DEV)
    with torch.no_grad():
        ops.head_fwd_bwd(pooled.detach(), labels, [w.detach() for w in ws], [b.detach() for b in bs], wc.detach(),
                         bc.detach(), l2, buf, train=True)
        dws = [torch.full_like(w, 7.0) for w in ws]          # overwritten, not accumulated
        dbs = [torch.full_like(b, 7.0) for b in bs]
        dwc, dbc = torch.full_like(wc, 7.0), torch.full_like(bc, 7.0)
        got_loss = ops.head_wgrad(pooled.detach(), buf, dws, dbs, dwc, dbc)
    assert _rel(buf.logits, logits) <= 1e-5
    assert abs(float(got_loss) - float(loss)) <= 1e-5 * abs(float(loss))
    assert _rel(buf.dpooled, pooled.grad) <= 1e-4
    for l in range(L):
        assert _rel(dws[l], ws[l].grad) <= 1e-4, l
        assert _rel(dbs[l], bs[l].grad) <= 1e-4, l
    assert _rel(dwc, wc.grad) <= 1e-4
    assert _rel(dbc, bc.grad) <= 1e-4


def test_k6_eval_mode_writes_logits_and_loss_only():
    B, H, C = 20, 200, 42
    pooled = torch.randn(B, 3 * H, device=DEV)
    labels = torch.randint(0, C, (B,), device=DEV)
    w = [torch.randn(H, 3 * H, device=DEV) * 0.05, torch.randn(H, H, device=DEV) * 0.05]
    b = [torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)]
    wc, bc = torch.randn(C, H, device=DEV) * 0.1, torch.zeros(C, device=DEV)
    buf = ops.HeadBuffers(B, H, C, 2, DEV)
    buf.dpooled.fill_(123.0)
    ops.head_fwd_bwd(pooled, labels, w, b, wc, bc, 0.0, buf, train=False)
    h = F.relu(F.linear(F.relu(F.linear(pooled, w[0], b[0])), w[1], b[1]))
    ref = F.linear(h, wc, bc)
    assert _rel(buf.logits, ref) <= 1e-5
    assert abs(float(buf.loss_rows.sum()) - float(F.cross_entropy(ref, labels))) <= 1e-5
    assert float(buf.dpooled.min()) == 123.0


def test_k6_unsupported_shapes_are_refused():
    from gcn_over_pruned_trees_b200 import _lib
    B, H, C = 4, 30, 5          # H % 4 != 0
    buf = ops.HeadBuffers(B, H, C, 1, DEV)
    with pytest.raises(_lib.GptError):
        ops.head_fwd_bwd(torch.randn(B, 3 * H, device=DEV), torch.zeros(B, dtype=torch.int64, device=DEV),
                         [torch.randn(H, 3 * H, device=DEV)], [torch.zeros(H, device=DEV)],
                         torch.randn(C, H, device=DEV), torch.zeros(C, device=DEV), 0.0, buf)


# ---------------------------------------------------------------- K7 ------------------------------------------------

@pytest.mark.parametrize('n,scale_grads', [(281_000, 1.0), (281_003, 40.0), (5, 40.0), (3_000_000, 0.01)])
def test_k7_clip_sgd_matches_torch(n, scale_grads):
    torch.manual_seed(n)
    V, E, rows = 3000, 300, 700
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * scale_grads / np.sqrt(n)
    emb = torch.randn(V, E, device=DEV)
    words = torch.randint(0, V, (rows,), device=DEV)
    words[::13] = 0                                           # padding_idx rows never move
    topn = 2500                                               # rows >= topn are frozen (model/gcn.py:83-86)
    st = ops.SparseEmbeddingState(emb, topn)
    st.words = words
    live = torch.unique(words[(words != 0) & (words < topn)])
    st.G[live] = torch.randn(live.numel(), E, device=DEV) * scale_grads / 50
    for r in range(rows - 1, -1, -1):                         # owner = first token of every live word
        w = int(words[r])
        if w != 0 and w < topn:
            st.owner[w] = r
    # reference: clip_grad_norm_ + SGD over (flat, dense embedding gradient)
    p_ref, e_ref = p.clone().requires_grad_(), emb.clone().requires_grad_()
    p_ref.grad, e_ref.grad = g.clone(), st.G.clone()
    norm_ref = torch.nn.utils.clip_grad_norm_([p_ref, e_ref], 5.0)
    torch.optim.SGD([p_ref, e_ref], lr=0.3).step()

    partials = torch.zeros(1024, device=DEV)
    norm = torch.zeros((), device=DEV)
    counter = torch.tensor([11, 5], dtype=torch.int64, device=DEV)
    ops.update_sqnorm(g, st, partials)
    ops.update_apply(p, g, st, emb, partials, 5.0, 0.3, 1.0, norm, counter[1:])
    assert abs(float(norm) - float(norm_ref)) <= 1e-5 * float(norm_ref)
    assert _rel(p, p_ref) <= 1e-6
    assert _rel(emb, e_ref) <= 1e-6
    assert float(g.abs().max()) == 0.0 and float(st.G.abs().max()) == 0.0       # zero_grad is part of the kernel
    assert int(st.owner.min()) == 0x7fffffff
    assert counter.tolist() == [11, 6]


def test_k7_grad_scale_is_the_data_parallel_mean():
    n = 10_000
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV)
    p_ref = p.clone().requires_grad_()
    p_ref.grad = g / 4
    torch.nn.utils.clip_grad_norm_([p_ref], 5.0)
    torch.optim.SGD([p_ref], lr=0.3).step()
    partials = torch.zeros(1024, device=DEV)
    ops.update_sqnorm(g, None, partials)
    ops.update_apply(p, g, None, None, partials, 5.0, 0.3, 0.25)
    assert _rel(p, p_ref) <= 1e-6


# ---------------------------------------------------------------- engine ----------------------------------------------

def _autograd_grads(tr, batch):
    tr.fast_update = False                  # the per-op autograd Functions, not engine.FastUpdate
    tr.optimizer.zero_grad()
    loss = tr.update(batch)
    loss.backward()
    return loss.detach(), {n: p.grad.detach().clone() for n, p in tr.model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
@pytest.mark.parametrize('over', [dict(), dict(dataset='semeval', num_class=19, ner_dim=0),
                                  dict(dataset='semeval', num_class=10), dict(pooling='avg'),
                                  dict(mlp_layers=1, num_layers=3), dict(prune_k=-1, hidden_dim=64),
                                  dict(topn=300), dict(no_adj=True)])
def test_fused_step_gradients_match_autograd_path(over, gemm_mode):
    """Same weights, dropout off: FusedTrainStep's hand-ordered backward == autograd over the per-op Functions, in the
    FFMA mode and in the 3xTF32 tensor-core mode bench.py runs."""
    cfg = dict(vocab_size=900, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode=gemm_mode)
    cfg.update(over)
    opt = synth.tacred_opt(**cfg)
    torch.manual_seed(3)
    a = GCNTrainer(dict(opt))
    b = GCNTrainer(dict(opt))
    b.model.load_state_dict(a.model.state_dict())
    a.model.train()
    b.model.train()
    batch = synth.make_batch(21, batch_size=50, vocab_size=900, num_class=opt['num_class'], dataset=opt['dataset'])
    loss_ref, ref = _autograd_grads(a, batch)
    fused = FusedTrainStep(b)
    loss, logits, got = fused.gradients(batch)
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    for name, g in ref.items():
        assert name in got, name
        assert _rel(got[name], g) <= 1e-4, name
    assert set(got) == set(ref)


def test_fused_step_with_dropout_uses_the_same_streams_as_the_autograd_path():
    opt = synth.tacred_opt(vocab_size=900, cuda=True, gemm_mode='fp32')
    torch.manual_seed(4)
    a = GCNTrainer(dict(opt))
    b = GCNTrainer(dict(opt))
    b.model.load_state_dict(a.model.state_dict())
    b.model.gcn_model.gcn.rng_state.copy_(a.model.gcn_model.gcn.rng_state)
    a.model.train()
    b.model.train()
    batch = synth.make_batch(22, batch_size=50, vocab_size=900)
    fused = FusedTrainStep(b)               # advances the step word once, like the first autograd forward will
    loss_ref, ref = _autograd_grads(a, batch)
    loss, _, got = fused.gradients(batch)
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    for name, g in ref.items():
        assert _rel(got[name], g) <= 1e-4, name


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
def test_fused_steps_match_reference_call_sequence(gemm_mode):
    """N fused (graph-replayed) steps == N x [zero_grad, update, backward, clip_grad_norm_, SGD.step] (train.py:213-227),
    the latter on the per-op autograd path; also in the 3xTF32 mode bench.py runs."""
    over = dict(vocab_size=700, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode=gemm_mode)
    batches = [synth.make_batch(50 + i, batch_size=50, vocab_size=700, pad_to=64) for i in range(3)]

    def run(fused):
        torch.manual_seed(5)
        tr = GCNTrainer(synth.tacred_opt(**over))
        tr.fast_update = False
        tr.model.train()
        losses = []
        for step in range(9):
            bt = batches[step % 3]
            if fused:
                losses.append(float(tr.train_step(bt)))
            else:
                tr.optimizer.zero_grad()
                loss = tr.update(bt)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(tr.model.parameters(), tr.opt['max_grad_norm'])
                tr.optimizer.step()
                losses.append(loss.item())
        return losses, {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}, tr

    l0, s0, _ = run(False)
    l1, s1, tr = run(True)
    assert isinstance(tr._graphed, FusedTrainStep)
    assert tr._graphed.replays >= 5
    assert np.allclose(l0, l1, rtol=2e-5)
    for k in s0:
        assert _rel(s1[k], s0[k]) < 5e-5, k
    # the trainer API still works on the re-homed parameters: predict, save/load layout
    preds, probs, loss = tr.predict(batches[0])
    assert len(preds) == 50 and np.isfinite(loss)
    sd = tr.model.state_dict()
    assert sd['gcn_model.gcn.W.0.weight'].shape == (200, 360)


def test_autograd_graph_engine_is_still_selected_when_fused_cannot_run():
    tr = GCNTrainer(synth.tacred_opt(vocab_size=500, cuda=True, conv_l2=1e-4))
    assert FusedTrainStep.unsupported_reason(tr) is not None
    tr.model.train()
    batch = synth.make_batch(7, batch_size=50, vocab_size=500)
    losses = [float(tr.train_step(batch)) for _ in range(6)]
    assert isinstance(tr._graphed, GraphedTrainStep) and np.isfinite(losses).all()


def test_fused_step_lr_change_recaptures():
    tr = GCNTrainer(synth.tacred_opt(vocab_size=500, cuda=True, input_dropout=0.0, gcn_dropout=0.0))
    tr.model.train()
    batch = synth.make_batch(8, batch_size=50, vocab_size=500)
    for _ in range(4):
        tr.train_step(batch)
    w0 = tr.model.classifier.weight.detach().clone()
    tr.update_lr(0.0)
    tr.train_step(batch)
    assert torch.equal(tr.model.classifier.weight, w0)         # lr 0: nothing moves


# ---------------------------------------------------------------- K3 wgrad over live rows -----------------------------

@pytest.mark.parametrize('M,N,K,frac', [(2750, 200, 360, 0.25), (2750, 200, 200, 0.4), (50, 64, 85, 1.0),
                                        (5000, 30, 17, 0.02), (9000, 512, 360, 0.7), (300, 200, 360, 0.0)])
def test_k3_wgrad_over_live_rows(M, N, K, frac):
    g = torch.Generator().manual_seed(M + N)
    flags = (torch.rand(M, generator=g) < frac).to(torch.uint8).to(DEV) * 5
    dy = torch.randn(M, N, generator=g).to(DEV)
    x = torch.randn(M, K, generator=g).to(DEV)
    live = flags != 0
    ref = (dy * live[:, None]).double().t() @ x.double()
    dw = torch.full((N, K), 2.0, device=DEV)                  # accumulates into what is there
    ops.linear_wgrad(dy, x, out=dw, accumulate=True, flags=flags)     # rows with flags == 0 are never read
    assert _rel(dw - 2.0, ref) <= 1e-5 if frac > 0 else float((dw - 2.0).abs().max()) == 0.0
    dw2 = torch.zeros((N, K), device=DEV)
    ops.linear_wgrad(dy, x, out=dw2, accumulate=True, flags=None)     # no flags: every row
    assert _rel(dw2, dy.double().t() @ x.double()) <= 1e-5


@pytest.mark.parametrize('M,N,K,frac', [(40000, 512, 360, 1.0), (33000, 200, 360, 0.4), (65536 + 5, 512, 512, 0.7),
                                        (32768, 64, 100, 1.0), (50000, 200, 200, 0.0),
                                        # CTA pairs (cta_group::2, K > 416): odd slice count, K not a multiple of 64, ragged M
                                        (40001, 320, 512, 0.5), (33333, 200, 420, 1.0), (70000, 512, 448, 0.3)])
def test_k3_wgrad_on_tensor_cores_is_fp32_grade(M, N, K, frac):
    """csrc/wgrad_tcgen05.cu (MN-major operands straight from HBM, 3xTF32): dw += dy^T x over the live rows, against
    fp64 at the fp32 parity tolerance (1e-5 relative); rows with flags == 0 may hold anything.  (>= 2 n-slices: the CTA-pair
    form; the single-CTA form of the same shapes runs in a subprocess below -- the switch is read once per process.)"""
    assert ops.wgrad_tc_ok(M, N, K)
    g = torch.Generator().manual_seed(M + N + K)
    flags = None
    dy = torch.randn(M, N, generator=g).to(DEV)
    x = torch.randn(M, K, generator=g).to(DEV)
    if frac < 1.0:
        flags = (torch.rand(M, generator=g) < frac).to(torch.uint8).to(DEV) * 3
        dead = flags == 0
        dy[dead] = float('nan')                       # never accumulated
        x[dead] = float('inf')
    live = torch.ones(M, dtype=torch.bool, device=DEV) if flags is None else flags != 0
    ref = torch.where(live[:, None], dy, torch.zeros_like(dy)).double().t() @ \
        torch.where(live[:, None], x, torch.zeros_like(x)).double()
    dw = torch.full((N, K), 0.5, device=DEV)          # accumulates into what is there
    ops.linear_wgrad(dy, x, 'tf32x3', out=dw, accumulate=True, flags=flags)
    if frac > 0:
        assert _rel(dw - 0.5, ref) <= 1e-5
    else:
        assert float((dw - 0.5).abs().max()) == 0.0


# ---------------------------------------------------------------- K2 (last layer) + K4 in one launch ------------------

@pytest.mark.parametrize('B,T,H,k,seed', [(50, 64, 200, 1, 0), (50, 96, 200, -1, 1), (7, 33, 64, 0, 2), (20, 40, 100, 2, 3),
                                          (16, 24, 36, 1, 4)])
def test_k2_last_layer_with_fused_max_pooling_equals_k2_then_k4(B, T, H, k, seed):
    """gpt_gcn_aggregate_fwd_pool == gpt_gcn_aggregate_fwd followed by gpt_pool3_fwd(max): layer output, activation
    mask, pooled values and argmax rows bit for bit (ties, empty pools and entity tokens outside the tree included)."""
    assert ops.aggregate_pool_ok(B, T, H)
    batch = synth.make_batch(900 + seed, batch_size=B, vocab_size=500, pad_to=T, max_len=T, mean_len=min(36, T // 2))
    dev = [t.to(DEV) for t in batch[:8]]
    csr = ops.prune_csr(dev[5], dev[6], dev[7], dev[4], dev[1], k)
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(B * T, H, generator=g).to(DEV)
    y[torch.rand(B * T, generator=g).to(DEV) < 0.1] = 0.0          # whole rows of zeros: ties at the ReLU floor
    bias = (torch.randn(H, generator=g) * 0.1).to(DEV)
    out_ref, act_ref = ops.aggregate_fwd(y, csr, bias, want_act=True)
    pooled_ref, arg_ref = ops.pool3_fwd(out_ref, csr, ops.POOL_TYPES['max'])
    pooled, argmax, act, out = ops.aggregate_fwd_pool(y, csr, bias, want_out=True)
    assert torch.equal(out, out_ref) and torch.equal(act, act_ref)
    assert torch.equal(pooled, pooled_ref)
    assert torch.equal(argmax, arg_ref)
    pooled2, argmax2, act2, none = ops.aggregate_fwd_pool(y, csr, bias)       # without storing the layer output
    assert none is None and torch.equal(pooled2, pooled_ref) and torch.equal(argmax2, arg_ref) and torch.equal(act2, act_ref)


@pytest.mark.parametrize('B,T,H,k,seed', [(50, 64, 200, 1, 0), (50, 96, 200, -1, 1), (7, 33, 64, 0, 2), (20, 40, 100, 2, 3),
                                          (6, 512, 96, -1, 5)])
def test_k2_last_layer_backward_with_fused_pool_backward_equals_k4_then_k2(B, T, H, k, seed):
    """gpt_gcn_aggregate_bwd_pool == gpt_pool3_bwd_masked + gpt_gcn_aggregate_bwd_pre, bit for bit (dy and dbias),
    including pools that share an argmax row, empty pools (argmax -1) and the large-tile configuration."""
    batch = synth.make_batch(950 + seed, batch_size=B, vocab_size=500, pad_to=T, max_len=T, mean_len=min(36, T // 2))
    dev = [t.to(DEV) for t in batch[:8]]
    csr = ops.prune_csr(dev[5], dev[6], dev[7], dev[4], dev[1], k)
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(B * T, H, generator=g).to(DEV)
    bias = (torch.randn(H, generator=g) * 0.1).to(DEV)
    out, act = ops.aggregate_fwd(y, csr, bias, want_act=True)
    pooled, argmax = ops.pool3_fwd(out, csr, ops.POOL_TYPES['max'])
    argmax[0, :5] = -1                                              # an empty pool
    argmax[1 % B, H:H + 7] = argmax[1 % B, 0:7]                     # two pools meeting in one row
    dpooled = torch.randn(B, 3 * H, generator=g).to(DEV)
    db_ref = torch.zeros(H, device=DEV)
    gg = ops.pool3_bwd_masked(dpooled, argmax, csr, ops.POOL_TYPES['max'], H, act, 0.0)
    dy_ref = ops.aggregate_bwd_pre(gg, csr, dbias_out=db_ref)
    db = torch.zeros(H, device=DEV)
    dy = ops.aggregate_bwd_pool(dpooled, argmax, act, csr, H, dbias_out=db)
    assert torch.equal(dy.view(-1), dy_ref.view(-1))
    assert float((db - db_ref).abs().max()) <= 1e-5 * max(1.0, float(db_ref.abs().max()))   # atomics: order only


def test_fused_pooling_is_declined_for_large_sentence_tiles():
    assert not ops.aggregate_pool_ok(64, 512, 512)


# ---------------------------------------------------------------- K8 (virtual ranks on one GPU) ------------------------

def _sparse_state(emb, words, topn, scale, gen):
    V, E = emb.shape
    st = ops.SparseEmbeddingState(emb, topn)
    st.words = words
    live = torch.unique(words[(words != 0) & (words < topn)])
    st.G[live] = (torch.randn(live.numel(), E, generator=gen) * scale).to(DEV)
    first = {}
    for r, w in enumerate(words.tolist()):
        if w != 0 and w < topn and w not in first:
            first[w] = r
    for w, r in first.items():
        st.owner[w] = r
    return st


@pytest.mark.parametrize('W', (1, 2, 3, 8))
def test_k8_exchange_of_virtual_ranks_is_the_mean_gradient_and_bit_identical(W):
    """W exchange regions in ONE process stand in for W GPUs (peer pointers are ordinary device pointers here): after
    push x W, reduce + apply on every 'rank' must equal clip_grad_norm_ + SGD on the mean gradient, and the replicas
    must stay bit-identical.  Three steps, so both parities and all resets are exercised."""
    gen = torch.Generator().manual_seed(100 + W)
    n, V, E, cap, topn = 70_016, 900, 300, 512, 850
    regions = [ops.ExchangeRegion(W, cap, E, V, n) for _ in range(W)]
    ptrs = [r.ptr for r in regions]
    shape = regions[0].shape
    p0 = torch.randn(n, generator=gen).to(DEV)
    e0 = torch.randn(V, E, generator=gen).to(DEV)
    params = [p0.clone() for _ in range(W)]
    embs = [e0.clone() for _ in range(W)]
    p_ref, e_ref = p0.clone().requires_grad_(), e0.clone().requires_grad_()
    sgd = torch.optim.SGD([p_ref, e_ref], lr=0.3)
    partials = torch.zeros(max(1024, regions[0].n_partials), device=DEV)
    counter = torch.tensor([1, 0], dtype=torch.int64, device=DEV)
    states = [None] * W
    for step in range(3):
        grads, dense_G = [], []
        for r in range(W):
            n_rows = int(torch.randint(100, cap, (1,), generator=gen))
            words = torch.randint(0, V, (n_rows,), generator=gen).to(DEV)    # overlapping vocabularies
            g = (torch.randn(n, generator=gen) * (30.0 if step == 1 else 0.01)).to(DEV)   # step 1 clips
            if states[r] is None:
                states[r] = _sparse_state(embs[r], words, topn, 0.05, gen)
            else:                       # G / owner were cleaned by the previous push: refill
                st = _sparse_state(embs[r], words, topn, 0.05, gen)
                assert float(states[r].G.abs().max()) == 0.0 and int(states[r].owner.min()) == 0x7fffffff
                states[r] = st
            grads.append(g.clone())
            dense_G.append(states[r].G.clone())
            ops.dp_push(ptrs, r, shape, g, states[r])
            ops.dp_signal(ptrs, r, shape)
        for r in range(W):
            g_buf = torch.empty(n, device=DEV)
            ops.dp_reduce(ptrs, r, shape, g_buf, partials, signal=False)
            ops.dp_apply(ptrs[r], shape, params[r], g_buf, embs[r], partials, 5.0, 0.3, None, counter[1:])
            assert float(g_buf.abs().max()) == 0.0
        p_ref.grad = torch.stack(grads).sum(0) / W
        e_ref.grad = torch.stack(dense_G).sum(0) / W
        torch.nn.utils.clip_grad_norm_([p_ref, e_ref], 5.0)
        sgd.step()
        for r in range(W):
            assert torch.equal(params[r], params[0]) and torch.equal(embs[r], embs[0]), (step, r)
        assert _rel(params[0], p_ref) <= 2e-6, step
        assert _rel(embs[0], e_ref) <= 2e-6, step
    assert int(counter[1]) == 3 * W
    for r in regions:
        r.free()


def test_fused_step_through_the_exchange_path_world_1():
    """data_parallel=True with a single rank: K8 (push to self, reduce, apply) must train like K7."""
    over = dict(vocab_size=700, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode='fp32')
    batches = [synth.make_batch(60 + i, batch_size=50, vocab_size=700, pad_to=64) for i in range(2)]

    def run(dp):
        torch.manual_seed(6)
        tr = GCNTrainer(synth.tacred_opt(**over))
        tr.model.train()
        eng = FusedTrainStep(tr, data_parallel=dp, max_rows=4096)
        losses = [float(eng(batches[i % 2])) for i in range(7)]
        return losses, {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}

    l0, s0 = run(False)
    l1, s1 = run(True)
    assert np.allclose(l0, l1, rtol=1e-5)
    for k in s0:
        assert _rel(s1[k], s0[k]) < 1e-5, k


# ---------------------------------------------------------------- fused K2-backward prologue --------------------------

@pytest.mark.parametrize('T,H,k', [(64, 200, 1), (96, 64, -1), (512, 512, -1), (40, 100, 2)])
def test_k2_backward_prologue_fused_into_its_producers(T, H, k):
    """pool3_bwd_masked / linear_dgrad_masked + aggregate_bwd_pre == pool3_bwd / linear_dgrad + aggregate_bwd."""
    B = 6 if T == 512 else 40
    batch = synth.make_batch(91, batch_size=B, vocab_size=500, pad_to=T, max_len=T, mean_len=T // 2)
    dev = [t.to(DEV) if torch.is_tensor(t) else t for t in batch]
    csr = ops.prune_csr(dev[5], dev[6], dev[7], dev[4], dev[1], k)
    y = torch.randn(B * T, H, device=DEV)
    bias = torch.randn(H, device=DEV) * 0.1
    rng = torch.tensor([77, 3], dtype=torch.int64, device=DEV)
    for p in (0.0, 0.5, 0.3):
        out, act = ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, subseq=1, want_act=True)
        # producer 1: K4 backward
        pooled, argmax = ops.pool3_fwd(out, csr, 0)
        gp = torch.randn(B, 3 * H, device=DEV)
        dh = ops.pool3_bwd(gp, argmax, csr, 0, H)
        dy_ref, db_ref = ops.aggregate_bwd(dh, None, csr, drop_p=p, act=act)
        g = ops.pool3_bwd_masked(gp, argmax, csr, 0, H, act, p)
        db = torch.zeros(H, device=DEV)
        dy = ops.aggregate_bwd_pre(g, csr, dbias_out=db)
        assert torch.equal(dy, dy_ref)
        assert _rel(db, db_ref) <= 1e-5
        # producer 2: the dgrad GEMM epilogue (3xTF32), N_next -> H
        N2 = 200 if H != 512 else 512
        w = torch.randn(N2, H, device=DEV) / np.sqrt(H)
        ws = ops.weight_prep(w, 'tf32x3')
        dnext = torch.randn(B * T, N2, device=DEV)
        dx = ops.linear_dgrad(dnext, w, 'tf32x3', ws).view(B, T, H)
        dy_ref, db_ref = ops.aggregate_bwd(dx, None, csr, drop_p=p, act=act)
        g = ops.linear_dgrad_masked(dnext, w, ws, act, csr, p)
        assert g is not None
        db = torch.zeros(H, device=DEV)
        dy = ops.aggregate_bwd_pre(g.view(B, T, H), csr, dbias_out=db)
        assert torch.equal(dy, dy_ref)
        assert _rel(db, db_ref) <= 1e-5


def test_tensor_core_weight_gradient_over_compacted_rows_at_tacred_size(monkeypatch):
    """GPT_TC_WGRAD_SMALL=1 (off by default): the layers' weight gradients from ops.LiveRows-compacted rows
    on the tensor cores == the FFMA kernel over the live rows, 1e-5 relative, and the fused step still trains."""
    grads = []
    for on in ('0', '1'):
        monkeypatch.setenv('GPT_TC_WGRAD_SMALL', on)
        torch.manual_seed(2)
        tr = GCNTrainer(synth.tacred_opt(vocab_size=900, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode='tf32x3'))
        tr.model.train()
        batch = tuple(t.cuda() if torch.is_tensor(t) else t for t in synth.make_batch(31, batch_size=50, vocab_size=900))
        eng = FusedTrainStep(tr, capture=False)
        inputs, labels = batch[:8], batch[8]
        eng._backward(eng._forward(inputs, labels))
        torch.cuda.synchronize()
        gcn = tr.model.gcn_model.gcn
        grads.append([eng.flat.g(lin.weight).clone() for lin in gcn.W])
    for a, b in zip(*grads):
        assert float((a - b).abs().max() / a.abs().max()) <= 1e-5


def test_k2_backward_also_stores_the_live_rows_compactly():
    """gpt_gcn_aggregate_bwd_{pre,pool}_c: dy unchanged, and dy's live rows (gpt_live_rows) also at their compact position."""
    B, T, H = 23, 61, 200
    batch = synth.make_batch(9, batch_size=B, vocab_size=500, pad_to=T)
    dev = [t.cuda() if torch.is_tensor(t) else t for t in batch]
    csr = ops.prune_csr(dev[5], dev[6], dev[7], dev[4], dev[1], 1)
    live = ops.LiveRows(csr.flags)
    cnt = int(live.count)
    idx = csr.flags.view(-1).nonzero().flatten()
    g = torch.randn(B, T, H, device=DEV) * csr.flags.view(B, T, 1).ne(0)
    ref = ops.aggregate_bwd_pre(g, csr)
    dyc = torch.full((B * T, H), float('nan'), device=DEV)
    got = ops.aggregate_bwd_pre(g, csr, live=live, compact_out=dyc)
    assert torch.equal(got, ref) and torch.equal(dyc[:cnt], ref[idx]) and bool(dyc[cnt:].isnan().all())
    # the pooled variant: act bits and argmax rows from a forward of the last layer
    y = torch.randn(B * T, H, device=DEV)
    pooled, argmax, act, _ = ops.aggregate_fwd_pool(y, csr, torch.zeros(H, device=DEV))
    dpooled = torch.randn(B, 3 * H, device=DEV)
    ref = ops.aggregate_bwd_pool(dpooled, argmax, act, csr, H)
    dyc.fill_(float('nan'))
    got = ops.aggregate_bwd_pool(dpooled, argmax, act, csr, H, live=live, compact_out=dyc)
    assert torch.equal(got, ref) and torch.equal(dyc[:cnt], ref[idx]) and bool(dyc[cnt:].isnan().all())


def test_k3_wgrad_single_cta_form_in_a_subprocess():
    """GPT_WGRAD_PAIR=0 (read once per process): the single-CTA form of the weight gradient on the shapes that take CTA
    pairs by default (odd slice counts, K not a multiple of 64, ragged M), against float64."""
    import os
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from gcn_over_pruned_trees_b200 import ops\n"
        "for M, N, K, frac in [(40001, 320, 512, 0.5), (33333, 200, 420, 1.0), (70000, 512, 448, 0.3), (40000, 512, 360, 1.0)]:\n"
        "    g = torch.Generator().manual_seed(M)\n"
        "    dy = torch.randn(M, N, generator=g).cuda(); x = torch.randn(M, K, generator=g).cuda()\n"
        "    flags = (torch.rand(M, generator=g) < frac).to(torch.uint8).cuda() * 3\n"
        "    live = (flags != 0)[:, None]\n"
        "    ref = (dy * live).double().t() @ (x * live).double()\n"
        "    dy[~live[:, 0]] = float('nan')\n"
        "    dw = torch.zeros(N, K, device='cuda')\n"
        "    ops.linear_wgrad(dy, x, 'tf32x3', out=dw, accumulate=True, flags=flags)\n"
        "    rel = float((dw.double() - ref).abs().max() / ref.abs().max())\n"
        "    assert rel <= 1e-5, (M, N, K, rel)\n"
        "print('SINGLE OK')\n")
    env = dict(os.environ, GPT_WGRAD_PAIR='0',
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'SINGLE OK' in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize('drop_p', (0.0, 0.5))
def test_front_kernel_equals_embedding_forward_plus_weight_preparation(drop_p):
    """gpt_embed_fwd_prep (K5 forward + the 3xTF32 operand preparation of every layer in one launch) == gpt_embed_fwd and
    gpt_weight_prep_tf32x3_batch, bit for bit (same Philox stream, same rounding)."""
    g = torch.Generator(device=DEV).manual_seed(4)
    B, T, V = 50, 67, 3000
    words = torch.randint(0, V, (B, T), device=DEV, generator=g)
    pos = torch.randint(0, 40, (B, T), device=DEV, generator=g)
    ner = torch.randint(0, 20, (B, T), device=DEV, generator=g)
    emb, pw, nw = (torch.randn(n, d, device=DEV, generator=g) for n, d in ((V, 300), (47, 30), (24, 30)))
    weights = [torch.randn(200, 360, device=DEV, generator=g), torch.randn(200, 200, device=DEV, generator=g),
               torch.randn(36, 100, device=DEV, generator=g)]
    rng = torch.tensor([12345, 7], dtype=torch.int64, device=DEV)
    ref_x = ops.embed_fwd(words, pos, ner, emb, pw, nw, drop_p, rng, 0xE0)
    ref_ws = [ops.weight_prep_buffer(w, 'tf32x3') for w in weights]
    ops.weight_prep_all(weights, 'tf32x3', ref_ws)
    outs = [torch.full_like(o, float('nan')) for o in ref_ws]
    x = ops.embed_fwd_prep(words, pos, ner, emb, pw, nw, drop_p, rng, 0xE0, weights, outs)
    torch.cuda.synchronize()
    assert torch.equal(x, ref_x)
    for a, b in zip(outs, ref_ws):
        assert torch.equal(a, b)
    assert ops.embed_fwd_prep(words, pos, ner, emb, pw, nw, drop_p, rng, 0xE0, weights, [None] + outs[1:]) is None
