"""GPU parity tests of the relation-aware adjacency modes (K10, csrc/deprel.cu): the whole model
with adj_type 'full_deprel' / 'diagonal_deprel' on the device, through the C ABI, nothing replaced -- against the real
reference's outputs (tests/golden/deprel.npz) and against the oracle with identical injected masks.

Tolerances: logits / pooled output / loss <= 1e-5 relative (max|d| / max|ref|), gradients <= 1e-4 relative per tensor,
in both projection modes ('fp32' FFMA and 'tf32x3' tcgen05).
"""
import numpy as np
import pytest
import torch

import cases
import weights
from gcn_over_pruned_trees_b200 import ops, synth
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
from oracle import gcn_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda'
_ALL = dict(cases.DEPREL_CASES, **cases.DEPREL_RANDOM_CASES)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _setup(golden_adj, name, gemm_mode='fp32', batch_size=None, table=None):
    over, source, wseed = (_ALL if table is None else table)[name]
    over = dict(over, gemm_mode=gemm_mode)
    if source[0] == 'split':
        batch = cases.batch_from_npz(golden_adj, source[1])
        over = dict(over, vocab_size=int(golden_adj['vocab_size']))
    else:
        batch = cases.make_case_batch((source[0], source[1], batch_size or source[2]), over)
    opt = synth.tacred_opt(**over)
    state = {k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()}
    trainer = GCNTrainer(dict(opt, cuda=True))
    trainer.model.load_state_dict(state)
    oracle = gcn_oracle.DenseClassifier(opt)
    oracle.load_state_dict(state)
    return opt, batch, trainer, oracle


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
@pytest.mark.parametrize('name', sorted(cases.DEPREL_CASES))
def test_relation_modes_eval_match_reference_outputs(golden_adj, golden_deprel, name, gemm_mode):
    opt, batch, trainer, _ = _setup(golden_adj, name, gemm_mode)
    trainer.model.eval()
    with torch.no_grad():
        logits, h_out = trainer.model([t.to(DEV) for t in batch[:-2]])
        loss = trainer.update(batch)
    assert _rel(logits.cpu(), golden_deprel['%s/logits' % name]) <= 1e-5
    assert _rel(h_out.cpu(), golden_deprel['%s/h_out' % name]) <= 1e-5
    assert abs(loss.item() - float(golden_deprel['%s/eval_loss' % name])) <= 1e-5 * abs(loss.item())
    preds, probs, ploss = trainer.predict(batch)                     # unsorted back to the loader's original order
    in_batch_order = logits.argmax(1).cpu().tolist()
    assert preds == [p for _, p in sorted(zip(batch[-1], in_batch_order))]
    assert np.isfinite(ploss) and _rel(np.asarray(probs).sum(1), np.ones(len(preds))) <= 1e-5


def _injected_masks(opt, batch, seed, edges=False, forget=False):
    g = torch.Generator().manual_seed(seed)
    B, T = batch[0].shape
    width = opt['emb_dim'] + opt['pos_dim'] + (opt['ner_dim'] if opt['dataset'] == 'tacred' else 0)

    def bern(shape, p):
        return (torch.rand(shape, generator=g) < p).float()

    masks = {'in': bern((B, T, width), 0.5) * 2.0}
    if opt.get('rnn', False):
        masks['rnn'] = bern((B, T, 2 * opt['rnn_hidden']), 0.5) * 2.0
    for l in range(opt['num_layers'] - 1):
        masks['gcn%d' % l] = bern((B, T, opt['hidden_dim']), 0.5) * 2.0
    for l in range(opt['num_layers']):
        if edges:
            masks['edge_f%d' % l] = bern((B, T, T), 0.6)
            masks['edge_r%d' % l] = bern((B, T, T), 0.6)
        if forget:
            masks['forget_f%d' % l] = bern((B, T, 1), 0.5)
            masks['forget_r%d' % l] = bern((B, T, 1), 0.5)
    return masks


def _compare_grads(trainer, oracle, tol=1e-4):
    got = dict(trainer.model.named_parameters())
    checked = 0
    for key, p in oracle.named_parameters():
        if p.grad is None:
            assert got[key].grad is None or float(got[key].grad.abs().max()) == 0.0, key
            continue
        assert got[key].grad is not None, key
        assert _rel(got[key].grad.cpu(), p.grad) <= tol, (key, _rel(got[key].grad.cpu(), p.grad))
        checked += 1
    return checked


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
@pytest.mark.parametrize('name,edges,forget', [
    ('full_k1_d8', False, False), ('full_k1_d8', True, True), ('full_kfull_d16', True, False),
    ('full_directed', True, False), ('full_no_self_loop', False, True), ('full_depth1_3layer', True, True),
    ('full_cgcn_h64', False, False), ('full_semeval', True, True), ('full_split_train', False, False),
    ('diag_k1', False, False), ('diag_kfull_3layer', False, False), ('diag_cgcn', False, False),
    ('full_entities_outside_tree', True, False), ('diag_entities_outside_tree', False, False)])
def test_relation_modes_train_grads_match_oracle(golden_adj, name, edges, forget, gemm_mode):
    """Train mode, every random draw (dropouts, edge dropout, relation forgetting) injected into both sides."""
    opt, batch, trainer, oracle = _setup(golden_adj, name, gemm_mode)
    masks = _injected_masks(opt, batch, seed=len(name), edges=edges, forget=forget)
    oracle.train()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    trainer.model.train()
    trainer.model.gcn_model.gcn.injected_masks = {k: v.to(DEV) for k, v in masks.items()}
    loss = trainer.update(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle) >= 8
    assert float(trainer.model.gcn_model.deprel_emb.weight.grad[0].abs().max()) == 0.0      # padding_idx row


def test_rows_beyond_one_wave_of_ctas(golden_adj):
    """More token rows than the grid has CTAs (148 x 16): every CTA walks several rows (shared-memory reuse across
    rows, per-CTA partial sums of the self-loop vector's gradient)."""
    table = {'big': (dict(cases.DEPREL_CASES['full_k1_d8'][0], prune_k=-1), ('synth', 341, 72), 51)}
    opt, batch, trainer, oracle = _setup(golden_adj, 'big', 'tf32x3', table=table)
    assert batch[0].numel() > 148 * 16
    masks = _injected_masks(opt, batch, seed=3, edges=True, forget=True)
    oracle.train()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    trainer.model.train()
    trainer.model.gcn_model.gcn.injected_masks = {k: v.to(DEV) for k, v in masks.items()}
    loss = trainer.update(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle) >= 8


def test_in_kernel_edge_dropout_and_forgetting_equal_their_materialised_masks(golden_adj):
    """The Philox decisions taken inside agg3 / drawn by gpt_relation_keep_tokens, fed to the oracle as dense masks."""
    table = {'philox': (dict(_ALL['full_edge_drop'][0], deprel_keep_prop=0.5), ('synth', 331, 16), 43)}
    opt, batch, trainer, oracle = _setup(golden_adj, 'philox', table=table)
    gcn = trainer.model.gcn_model.gcn
    masks = _injected_masks(opt, batch, seed=5)              # dropouts injected; edges / forgetting left to Philox
    gcn.injected_masks = {k: v.to(DEV) for k, v in masks.items()}
    gcn.rng_state.copy_(torch.tensor([1234567, 3]))
    B, T = batch[0].shape
    for l in range(opt['num_layers']):
        masks['edge_f%d' % l] = ops.edge_keep_dense(gcn.rng_state, B, T, l, 0, opt['edge_keep_prob']).float().cpu()
        masks['edge_r%d' % l] = ops.edge_keep_dense(gcn.rng_state, B, T, l, 1, opt['edge_keep_prob']).float().cpu()
        kf, kr = ops.relation_keep_tokens(gcn.rng_state, B * T, l, opt['deprel_keep_prop'])
        masks['forget_f%d' % l], masks['forget_r%d' % l] = kf.view(B, T, 1).float().cpu(), kr.view(B, T, 1).float().cpu()
        for m, p in ((masks['edge_f%d' % l], 0.7), (masks['edge_r%d' % l], 0.7), (masks['forget_f%d' % l], 0.5)):
            assert abs(float(m.mean()) - p) < 0.06
        assert not torch.equal(masks['edge_f%d' % l], masks['edge_r%d' % l])
    trainer.model.train()
    oracle.train()
    loss = trainer.update(batch)
    loss.backward()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle) >= 8


def test_in_kernel_dropout_rate_and_backward_consistency(golden_adj):
    opt, batch, trainer, _ = _setup(golden_adj, 'full_k1_d8')
    gcn = trainer.model.gcn_model.gcn
    B, T = batch[0].shape
    H = opt['hidden_dim']
    deprel, head, subj_pos, obj_pos = [t.to(DEV) for t in batch[4:8]]
    csr = ops.prune_csr(head, subj_pos, obj_pos, deprel, batch[1].to(DEV), opt['prune_k'])
    g = torch.Generator(device=DEV).manual_seed(0)
    F, R, S = (torch.randn(B * T, H, device=DEV, generator=g) for _ in range(3))
    base = ops._agg3_fwd(F, R, S, csr, ops.RelationLayerConfig(0, rng_state=gcn.rng_state))
    cfg = ops.RelationLayerConfig(0, drop_p=0.5, rng_state=gcn.rng_state)
    out = ops._agg3_fwd(F, R, S, csr, cfg)
    live = base > 0
    kept = (out != 0) & live
    assert abs(float(kept.sum()) / float(live.sum()) - 0.5) < 0.03
    assert torch.allclose(out[kept], base[kept] * 2.0)
    gout = torch.randn(B, T, H, device=DEV, generator=g)
    dF, dR, dS = ops._agg3_bwd(gout, out, csr, cfg)
    cfg_m = ops.RelationLayerConfig(0, drop_mask=kept.float().view(B * T, H) * 2.0, rng_state=gcn.rng_state)
    dF2, dR2, dS2 = ops._agg3_bwd(gout, ops._agg3_fwd(F, R, S, csr, cfg_m), csr, cfg_m)
    for a, b in ((dF, dF2), (dR, dR2), (dS, dS2)):
        assert torch.equal(a, b)


@pytest.mark.parametrize('name', ('full_k1_d8', 'diag_k1', 'full_edge_drop', 'full_forget'))
def test_relation_modes_train_loop_lowers_the_loss(golden_adj, name):
    """train.py:213-227 on the new package (eager), every random feature drawn in-kernel."""
    opt, batch, trainer, _ = _setup(golden_adj, name, 'tf32x3')
    trainer.model.train()
    losses = []
    for _ in range(6):
        trainer.optimizer.zero_grad()
        loss = trainer.update(batch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), opt['max_grad_norm'])
        trainer.optimizer.step()
        losses.append(loss.item())
        del loss
    assert np.isfinite(losses).all() and min(losses[3:]) < losses[0]


@pytest.mark.parametrize('name', ('full_k1_d8', 'diag_k1', 'full_edge_drop', 'full_forget'))
def test_relation_modes_graphed_step_lowers_the_loss(golden_adj, name):
    """trainer.train_step: the same step captured into a CUDA graph (autograd under capture, GraphedTrainStep).  A
    fresh trainer: a capture that followed eager update()/backward() steps whose last loss tensor was still referenced
    failed once with cudaErrorStreamCaptureImplicit (suspected: AccumulateGrad nodes kept alive
    by the old graph carry the default stream)."""
    opt, batch, trainer, _ = _setup(golden_adj, name, 'tf32x3')
    trainer.model.train()
    graphed = [float(trainer.train_step(batch)) for _ in range(10)]
    assert type(trainer._graphed).__name__ == 'GraphedTrainStep' and trainer._graphed.replays >= 6
    assert np.isfinite(graphed).all() and min(graphed[5:]) < graphed[0]


def test_relation_mode_checkpoint_round_trip(golden_adj, tmp_path):
    from gcn_over_pruned_trees_b200 import torch_utils
    opt, batch, a, _ = _setup(golden_adj, 'full_k1_d8')
    f = str(tmp_path / 'ckpt.pt')
    a.save(f, 1)
    b = GCNTrainer(torch_utils.load_config(f))
    b.load(f)
    pa, _, la = a.predict(batch)
    pb, _, lb = b.predict(batch)
    assert pa == pb and la == lb
    sd = b.model.state_dict()
    assert sd['gcn_model.deprel_emb.weight'].data_ptr() == sd['gcn_model.gcn.deprel_emb.weight'].data_ptr()
    assert tuple(sd['gcn_model.gcn.W.weight'].shape) == (8 * 64, 64)


@pytest.mark.parametrize('frac', (0.0, 0.27, 1.0))
def test_live_row_compaction_and_row_limited_projections(frac):
    """gpt_live_rows / gather / scatter and the projections that take their row count from the device
    (gpt_linear_{fwd,dgrad,wgrad}_tf32x3_rows): the rows beyond the count hold NaN scratch here and must never leak.
    3xTF32 tolerance 1e-5 relative of the largest element, against float64."""
    g = torch.Generator(device=DEV).manual_seed(11)
    N, K, DH = 3000, 200, 2560
    flags = (torch.rand(N, device=DEV, generator=g) < frac).to(torch.uint8) * 3
    live = ops.LiveRows(flags)
    idx = flags.nonzero().flatten()
    cnt = int(live.count)
    assert cnt == idx.numel()
    assert torch.equal(live.perm[:cnt].long(), idx)
    expect_inv = torch.full((N,), -1, dtype=torch.int32, device=DEV)
    expect_inv[idx] = torch.arange(cnt, dtype=torch.int32, device=DEV)
    assert torch.equal(live.inv, expect_inv)
    assert torch.equal(live.live, (torch.arange(N, device=DEV) < cnt).to(torch.uint8))

    x = torch.randn(N, K, device=DEV, generator=g)
    xc = live.gather(x)
    assert torch.equal(xc[:cnt], x[idx])
    xc[cnt:] = float('nan')
    w = torch.randn(DH, K, device=DEV, generator=g) / np.sqrt(K)
    ws = ops.weight_prep(w, 'tf32x3')
    bias = torch.randn(DH, device=DEV, generator=g)
    y, biased = ops._linear_fwd_rows(xc, w, 'tf32x3', ws, live, bias)     # the bias rides in the projection's epilogue
    assert biased
    dy = torch.randn(N, DH, device=DEV, generator=g)
    dy[cnt:] = float('nan')
    dxc = ops._linear_dgrad_rows(dy, w, 'tf32x3', ws, live)
    dx = live.scatter(dxc)
    dw = ops._linear_wgrad_rows(dy, xc, 'tf32x3', live)
    torch.cuda.synchronize()
    dead = flags == 0
    assert not bool(dx[dead].ne(0).any()) and torch.equal(dx[idx], dxc[:cnt])
    if cnt == 0:
        assert not bool(dw.ne(0).any())
        return
    assert _rel(y[:cnt].cpu(), (xc[:cnt].double() @ w.double().t() + bias.double()).cpu()) <= 1e-5
    assert _rel(dxc[:cnt].cpu(), (dy[:cnt].double() @ w.double()).cpu()) <= 1e-5
    assert _rel(dw.cpu(), (dy[:cnt].double().t() @ xc[:cnt].double()).cpu()) <= 1e-5
    db = torch.zeros(DH, device=DEV)
    ops._call('gpt_colsum_acc_rows', dy.data_ptr(), N, DH, live.count.data_ptr(), db.data_ptr(), ops._stream())
    assert _rel(db.cpu(), dy[:cnt].double().sum(0).cpu()) <= 1e-5


def test_full_deprel_compacted_rows_equal_all_rows(golden_adj, monkeypatch):
    """The layer over the observable rows only == the layer over every row (GPT_K10_COMPACT=0): logits, loss and every
    gradient, 1e-5 / 1e-4 (the two differ in summation order only)."""
    res = []
    for compact in ('1', '0'):
        monkeypatch.setenv('GPT_K10_COMPACT', compact)
        opt, batch, trainer, _ = _setup(golden_adj, 'full_k1_d8', 'tf32x3')
        trainer.model.train()
        masks = _injected_masks(opt, batch, seed=5, edges=True, forget=True)
        trainer.model.gcn_model.gcn.injected_masks = {k: v.to(DEV) for k, v in masks.items()}
        loss = trainer.update(batch)
        loss.backward()
        res.append((loss.item(), {n: p.grad.detach().cpu().numpy() for n, p in trainer.model.named_parameters()
                                  if p.grad is not None}))
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[1][0])
    assert len(res[0][1]) == len(res[1][1]) >= 8
    for n, g in res[1][1].items():
        assert _rel(res[0][1][n], g) <= 1e-4, n
