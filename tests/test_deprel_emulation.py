"""CPU: the SOURCE of csrc/deprel.cu (K10, relation-aware layers), csrc/prune_csr.cu (K1), csrc/embed.cu (K5),
csrc/gemm_simt.cu (K3, fp32 mode) and csrc/pool3.cu (K4) executed on the host (tests/emu: one fiber per CUDA thread, barriers for __syncthreads / shuffles) underneath the
product's own Python layers -- model/gcn.py -> ops.py autograd Functions -> C ABI -- and checked against the real
reference's outputs (tests/golden/deprel.npz) and against the oracle with identical injected masks.

The build container has no GPU, so the pieces of the path that only exist as GPU code are replaced here, and only
here: cuDNN's LSTM runs as torch's CPU LSTM, and where dropout masks are injected the embedding stage is torch's
lookup (K5 draws its own Philox mask; the eval-mode cases run K5 itself).  What this file pins is therefore K1's CSR as
K10 and K4 consume it, the embedding stage, the projections and their gradients, K10's arithmetic, its direction / edge / forgetting conventions, the weight_l re-layout and
every backward formula; the `-m gpu` tests in test_gpu_relation_modes.py run the same cases on the device with nothing
replaced.
"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

import cases
import weights
from gcn_over_pruned_trees_b200 import _lib, ops, synth
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
from oracle import gcn_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))

K10 = ('gpt_prune_csr', 'gpt_pool3_fwd', 'gpt_pool3_bwd', 'gpt_linear_fwd_f32', 'gpt_linear_dgrad_f32',
       'gpt_linear_wgrad_f32', 'gpt_embed_fwd', 'gpt_embed_bwd', 'gpt_relmix_fwd', 'gpt_relmix_bwd', 'gpt_diagmix_fwd', 'gpt_diagmix_bwd', 'gpt_agg3_fwd', 'gpt_agg3_bwd',
       'gpt_edge_keep_dense', 'gpt_relation_keep_tokens', 'gpt_colsum_acc', 'gpt_live_rows', 'gpt_gather_rows', 'gpt_scatter_rows', 'gpt_relmix_fwd_rows', 'gpt_relmix_bwd_rows',
       'gpt_colsum_acc_rows', 'gpt_linear_wgrad_rows_f32')
_ALL = dict(cases.DEPREL_CASES, **cases.DEPREL_RANDOM_CASES)


@pytest.fixture(scope='module')
def emulated(request):
    import emu_build
    handle = ctypes.CDLL(emu_build.build())
    for name in K10:
        fn = getattr(handle, name)
        fn.argtypes = _lib.SIGNATURES[name]
        fn.restype = ctypes.c_int
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, '_lib', handle)
    mp.setattr(ops, '_dev', lambda t, dtype, name: t.contiguous() if t.dtype == dtype else (_ for _ in ()).throw(
        TypeError('%s must be %s' % (name, dtype))))
    mp.setattr(ops, '_stream', lambda: None)
    yield handle
    mp.undo()


def _setup(golden_adj, name, batch_size=None):
    over, source, wseed = _ALL[name]
    if source[0] == 'split':
        batch = cases.batch_from_npz(golden_adj, source[1])
        over = dict(over, vocab_size=int(golden_adj['vocab_size']))
    else:
        batch = cases.make_case_batch((source[0], source[1], batch_size or source[2]), over)
    opt = synth.tacred_opt(**dict(over, gemm_mode='fp32'))       # K3's FFMA kernels; tcgen05 cannot be emulated
    state = {k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()}
    trainer = GCNTrainer(dict(opt))
    trainer.model.load_state_dict(state)
    oracle = gcn_oracle.DenseClassifier(opt)
    oracle.load_state_dict(state)
    return opt, batch, trainer, oracle


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('name', sorted(cases.DEPREL_CASES))
def test_emulated_kernels_match_reference_eval(emulated, golden_adj, golden_deprel, name):
    opt, batch, trainer, _ = _setup(golden_adj, name)
    trainer.model.eval()
    with torch.no_grad():
        logits, h_out = trainer.model(list(batch[:-2]))
        loss = trainer.update(batch)
    assert _rel(logits, torch.from_numpy(golden_deprel['%s/logits' % name])) <= 1e-5
    assert _rel(h_out, torch.from_numpy(golden_deprel['%s/h_out' % name])) <= 1e-5
    assert abs(loss.item() - float(golden_deprel['%s/eval_loss' % name])) <= 1e-5 * abs(loss.item())


def _injected_masks(opt, batch, seed, edges=False, forget=False):
    g = torch.Generator().manual_seed(seed)
    B, T = batch[0].shape
    tacred = opt['dataset'] == 'tacred'
    width = opt['emb_dim'] + opt['pos_dim'] + (opt['ner_dim'] if tacred else 0)

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


def _compare_grads(trainer, oracle, tol=2e-5):
    got = dict(trainer.model.named_parameters())
    checked = 0
    for key, p in oracle.named_parameters():
        if p.grad is None:
            assert got[key].grad is None or float(got[key].grad.abs().max()) == 0.0, key
            continue
        assert got[key].grad is not None, key
        assert _rel(got[key].grad, p.grad) <= tol, (key, _rel(got[key].grad, p.grad))
        checked += 1
    return checked


@pytest.mark.parametrize('name,edges,forget', [
    ('full_k1_d8', False, False), ('full_k1_d8', True, True), ('full_directed', True, False),
    ('full_no_self_loop', False, True), ('full_depth1_3layer', True, True), ('full_cgcn_h64', False, False),
    ('full_semeval', True, True), ('diag_k1', False, False), ('diag_kfull_3layer', False, False),
    ('full_entities_outside_tree', True, False), ('diag_entities_outside_tree', False, False)])
def test_emulated_kernels_train_grads_match_oracle(emulated, golden_adj, name, edges, forget):
    """Train mode, every random draw injected into both sides: loss and all parameter gradients."""
    opt, batch, trainer, oracle = _setup(golden_adj, name, batch_size=None if 'outside_tree' in name else 8)
    masks = _injected_masks(opt, batch, seed=len(name), edges=edges, forget=forget)
    trainer.model.train()
    oracle.train()
    trainer.model.gcn_model.gcn.injected_masks = masks
    loss = trainer.update(batch)
    loss.backward()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle) >= 8
    g = trainer.model.gcn_model.deprel_emb.weight.grad
    assert float(g[0].abs().max()) == 0.0                    # padding_idx row


@pytest.mark.parametrize('name', ('full_k1_d8', 'diag_k1'))
def test_emulated_whole_model_train_step_without_dropout(emulated, golden_adj, name):
    """Nothing injected, nothing replaced (dropout probabilities 0): K1, K5 (forward + scatter backward), K3, K10 and K4
    from source under the product's autograd path; loss and every gradient against the oracle."""
    over, source, wseed = _ALL[name]
    _ALL['_nodrop'] = (dict(over, input_dropout=0.0, gcn_dropout=0.0), (source[0], source[1], 8), wseed)
    try:
        opt, batch, trainer, oracle = _setup(golden_adj, '_nodrop')
    finally:
        del _ALL['_nodrop']
    assert trainer.model.gcn_model.gcn.injected_masks is None
    trainer.model.train()
    oracle.train()
    loss = trainer.update(batch)
    loss.backward()
    ref_loss, _ = oracle.loss(batch)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle) >= 8


def test_emulated_rows_beyond_one_wave_of_ctas(emulated, golden_adj, monkeypatch):
    """Fewer CTAs than token rows (GPT_K10_MAX_CTAS; on the device the cap is 148 x 16): every CTA walks several rows --
    shared-memory reuse across rows, per-CTA partial sums of the self-loop vector's gradient."""
    monkeypatch.setenv('GPT_K10_MAX_CTAS', '24')
    opt, batch, trainer, oracle = _setup(golden_adj, 'full_kfull_d16', batch_size=8)
    assert batch[0].numel() > 24 * 8
    masks = _injected_masks(opt, batch, seed=3, edges=True, forget=True)
    trainer.model.train()
    oracle.train()
    trainer.model.gcn_model.gcn.injected_masks = masks
    loss = trainer.update(batch)
    loss.backward()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle) >= 8


def test_in_kernel_edge_dropout_and_forgetting_equal_their_materialised_masks(emulated, golden_adj):
    """The Philox decisions taken inside agg3 / drawn by gpt_relation_keep_tokens, fed to the oracle as dense masks."""
    name = 'full_edge_drop'
    over = dict(_ALL[name][0], deprel_keep_prop=0.5)
    _ALL['_philox'] = (over, ('synth', 331, 8), 43)
    try:
        opt, batch, trainer, oracle = _setup(golden_adj, '_philox')
    finally:
        del _ALL['_philox']
    gcn = trainer.model.gcn_model.gcn
    masks = _injected_masks(opt, batch, seed=5)              # dropouts injected; edges / forgetting left to Philox
    gcn.injected_masks = dict(masks)
    gcn.rng_state[0], gcn.rng_state[1] = 1234567, 3
    B, T = batch[0].shape
    for l in range(opt['num_layers']):
        masks['edge_f%d' % l] = ops.edge_keep_dense(gcn.rng_state, B, T, l, 0, opt['edge_keep_prob']).float()
        masks['edge_r%d' % l] = ops.edge_keep_dense(gcn.rng_state, B, T, l, 1, opt['edge_keep_prob']).float()
        kf, kr = ops.relation_keep_tokens(gcn.rng_state, B * T, l, opt['deprel_keep_prop'])
        masks['forget_f%d' % l], masks['forget_r%d' % l] = kf.view(B, T, 1).float(), kr.view(B, T, 1).float()
        for m, p in ((masks['edge_f%d' % l], 0.7), (masks['edge_r%d' % l], 0.7), (masks['forget_f%d' % l], 0.5)):
            assert abs(float(m.mean()) - p) < 0.06
        assert not torch.equal(masks['edge_f%d' % l], masks['edge_r%d' % l])
    assert not torch.equal(masks['edge_f0'], masks['edge_f1'])
    trainer.model.train()
    oracle.train()
    loss = trainer.update(batch)
    loss.backward()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle) >= 8


def test_in_kernel_dropout_rate_and_backward_consistency(emulated, golden_adj):
    """gcn_drop drawn inside agg3: keep rate ~ 1-p, kept elements scaled by 1/(1-p), and the backward (which recovers
    d out / d z from out != 0) equals autograd through the same realised mask."""
    opt, batch, trainer, oracle = _setup(golden_adj, 'full_k1_d8', batch_size=8)
    gcn = trainer.model.gcn_model.gcn
    B, T = batch[0].shape
    H = opt['hidden_dim']
    inputs = list(batch[:-2])
    csr = ops.prune_csr(inputs[5], inputs[6], inputs[7], inputs[4], inputs[1], opt['prune_k'])
    g = torch.Generator().manual_seed(0)
    F, R, S = (torch.randn(B * T, H, generator=g) for _ in range(3))
    cfg0 = ops.RelationLayerConfig(0, rng_state=gcn.rng_state)
    cfg = ops.RelationLayerConfig(0, drop_p=0.5, rng_state=gcn.rng_state)
    base = ops._agg3_fwd(F, R, S, csr, cfg0)
    out = ops._agg3_fwd(F, R, S, csr, cfg)
    live = base > 0
    kept = (out != 0) & live
    assert abs(float(kept.sum()) / float(live.sum()) - 0.5) < 0.03
    assert torch.allclose(out[kept], base[kept] * 2.0)
    gout = torch.randn(B, T, H, generator=g)
    dF, dR, dS = ops._agg3_bwd(gout, out, csr, cfg)
    cfg_m = ops.RelationLayerConfig(0, drop_mask=kept.float().view(B * T, H) * 2.0, rng_state=gcn.rng_state)
    dF2, dR2, dS2 = ops._agg3_bwd(gout, ops._agg3_fwd(F, R, S, csr, cfg_m), csr, cfg_m)
    for a, b in ((dF, dF2), (dR, dR2), (dS, dS2)):
        assert torch.equal(a, b)


@pytest.mark.parametrize('n_rows,frac', [(1, 1.0), (37, 0.0), (300, 0.3), (2500, 0.6)])
def test_emulated_live_row_compaction_gather_scatter_and_column_sums(emulated, n_rows, frac):
    """gpt_live_rows / gpt_gather_rows / gpt_scatter_rows / gpt_colsum_acc_rows from source on the CPU against numpy
    (several 1024-row rounds of the single-CTA scan, empty and full lists, K not a multiple of 4)."""
    g = torch.Generator().manual_seed(n_rows)
    flags = (torch.rand(n_rows, generator=g) < frac).to(torch.uint8) * 5
    live = ops.LiveRows(flags)
    idx = flags.nonzero().flatten()
    cnt = int(live.count)
    assert cnt == idx.numel()
    assert torch.equal(live.perm[:cnt].long(), idx)
    inv = torch.full((n_rows,), -1, dtype=torch.int32)
    inv[idx] = torch.arange(cnt, dtype=torch.int32)
    assert torch.equal(live.inv, inv)
    assert torch.equal(live.live, (torch.arange(n_rows) < cnt).to(torch.uint8))
    for K in (8, 7):
        x = torch.randn(n_rows, K, generator=g)
        xc = live.gather(x, out=torch.full((n_rows, K), 9.0))
        assert torch.equal(xc[:cnt], x[idx]) and bool((xc[cnt:] == 9.0).all())
        back = live.scatter(xc)
        assert torch.equal(back[idx], x[idx]) and not bool(back[flags == 0].ne(0).any())
        s = torch.zeros(K)
        ops._call('gpt_colsum_acc_rows', xc.data_ptr(), n_rows, K, live.count.data_ptr(), s.data_ptr(), None)
        assert torch.allclose(s, x[idx].sum(0), rtol=1e-5, atol=1e-5)
