"""GPU tests of the reference-facing training loop on captured graphs (engine.FastUpdate: update / backward / step of
train.py:213-227 as three CUDA-graph replays) and of the fused step exactly as bench.py runs it (3xTF32 + in-kernel
Philox dropout), against the per-op autograd path and the oracle.

Tolerances: gradients <= 1e-4 relative per tensor, parameters after N steps <= 5e-5, losses <= 2e-5.
"""
import numpy as np
import pytest
import torch

from gcn_over_pruned_trees_b200 import ops, synth
from gcn_over_pruned_trees_b200.engine import FastSGD, FastUpdate, FusedTrainStep, GraphedTrainStep
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
from oracle import gcn_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    a = a.detach().double().cpu().numpy()
    b = b.detach().double().cpu().numpy()
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _pair(over, seed=5):
    opt = synth.tacred_opt(**over)
    torch.manual_seed(seed)
    a = GCNTrainer(dict(opt))
    b = GCNTrainer(dict(opt))
    b.model.load_state_dict(a.model.state_dict())
    b.model.gcn_model.gcn.rng_state.copy_(a.model.gcn_model.gcn.rng_state)
    a.model.train()
    b.model.train()
    a.fast_update, b.fast_update = False, True
    return a, b


def _five_calls(tr, batch):
    tr.optimizer.zero_grad()
    loss = tr.update(batch)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(tr.model.parameters(), tr.opt['max_grad_norm'])
    tr.optimizer.step()
    return loss.item()


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
def test_fast_update_matches_the_autograd_path_over_the_reference_loop(gemm_mode):
    """9 x (zero_grad, update, backward, clip_grad_norm_, optimizer.step) -- train.py:213-227 verbatim -- with the three
    trainer calls replaying captured graphs == the same calls on the per-op autograd Functions."""
    over = dict(vocab_size=700, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode=gemm_mode)
    batches = [synth.make_batch(50 + i, batch_size=50, vocab_size=700, pad_to=64) for i in range(3)]
    a, b = _pair(over)
    la = [_five_calls(a, batches[s % 3]) for s in range(9)]
    lb = [_five_calls(b, batches[s % 3]) for s in range(9)]
    assert isinstance(b._fast, FastUpdate) and isinstance(b.optimizer, FastSGD) and a._fast is None
    assert b._fast.replays >= 3 * 5             # steps 4..9 of the single shape: forward, backward and apply replays
    assert np.allclose(la, lb, rtol=2e-5)
    sa, sb = a.model.state_dict(), b.model.state_dict()
    for k in sa:
        assert _rel(sb[k], sa[k]) < 5e-5, k
    # lr decay as train.py:340-343 does it reaches the captured update
    w0 = b.model.classifier.weight.detach().clone()
    b.update_lr(0.0)
    _five_calls(b, batches[0])
    assert torch.equal(b.model.classifier.weight, w0)


def test_fast_update_gradients_and_autograd_semantics():
    """update() alone touches no gradient; backward() ACCUMULATES grad_output x gradient into .grad; a stale loss is
    refused; accumulated gradients take torch's SGD step over the same buffers."""
    over = dict(vocab_size=600, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode='fp32')
    b1 = synth.make_batch(71, batch_size=50, vocab_size=600)
    b2 = synth.make_batch(72, batch_size=30, vocab_size=600)
    a, b = _pair(over, seed=9)
    for tr in (a, b):
        tr.optimizer.zero_grad()
        for _ in range(4):                      # past the warm-up: graphs are what runs below
            tr.update(b1)
            tr.update(b2)
    assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in b.model.parameters())
    for tr in (a, b):                           # two micro-batches, the second scaled, before one optimizer step
        tr.optimizer.zero_grad()
        tr.update(b1).backward()
        (tr.update(b2) * 0.5).backward()
    ga = {n: p.grad for n, p in a.model.named_parameters() if p.grad is not None}
    gb = {n: p.grad for n, p in b.model.named_parameters() if p.grad is not None}
    assert set(ga) == set(gb)
    for n in ga:
        assert _rel(gb[n], ga[n]) <= 1e-4, n
    for tr in (a, b):
        torch.nn.utils.clip_grad_norm_(tr.model.parameters(), 5.0)
        tr.optimizer.step()
        tr.optimizer.zero_grad()
    for (n, pa), (_, pb) in zip(a.model.named_parameters(), b.model.named_parameters()):
        assert _rel(pb, pa) < 1e-5, n
    assert float(b._fast.engine.flat.grad.abs().max()) == 0.0 and float(b._fast.engine.sparse.G.abs().max()) == 0.0
    stale = b.update(b1)
    b.update(b1)                                # overwrites the activations `stale` would need
    with pytest.raises(RuntimeError, match='overwritten'):
        stale.backward()
    # and it keeps training through the fast path afterwards, interleaved with train_step on the same trainer
    b.optimizer.zero_grad()
    losses = [_five_calls(b, b1) for _ in range(3)] + [float(b.train_step(b1)) for _ in range(4)] + [_five_calls(b, b1)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def _materialised_masks(trainer, batch, csr):
    """The dropout masks the fused step will draw at the CURRENT {seed, step} of the model, as float tensors already
    scaled by 1/(1-p): K5's (input dropout) from two forwards with and without dropout, K2's per layer from an
    all-ones projection (every observable row has a positive pre-activation there)."""
    opt = trainer.opt
    gm = trainer.model.gcn_model
    rng = gm.gcn.rng_state
    words, pos, ner = batch[0].to(DEV), batch[2].to(DEV), batch[3].to(DEV)
    x0 = ops.embed_fwd(words, pos, ner, gm.emb.weight.data, gm.pos_emb.weight.data, gm.ner_emb.weight.data, 0.0, rng, 0xE0)
    x1 = ops.embed_fwd(words, pos, ner, gm.emb.weight.data, gm.pos_emb.weight.data, gm.ner_emb.weight.data,
                       opt['input_dropout'], rng, 0xE0)
    scale = ops.drop_scale(opt['input_dropout'])
    masks = {'in': torch.where(x0 != 0, x1 / torch.where(x0 != 0, x0, torch.ones_like(x0)),
                               torch.full_like(x0, scale)).cpu()}
    B, T = words.shape
    H = opt['hidden_dim']
    ones = torch.ones(B * T, H, device=DEV)
    zero_bias = torch.zeros(H, device=DEV)
    for l in range(opt['num_layers'] - 1):
        o0 = ops.aggregate_fwd(ones, csr, zero_bias, True, 0.0, rng, l)
        o1 = ops.aggregate_fwd(ones, csr, zero_bias, True, opt['gcn_dropout'], rng, l)
        masks['gcn%d' % l] = torch.where(o0 > 0, o1 / torch.where(o0 > 0, o0, torch.ones_like(o0)),
                                         torch.zeros_like(o0)).cpu()
    return masks


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
def test_fused_step_with_philox_dropout_matches_the_oracle_with_the_same_masks(gemm_mode):
    """The configuration bench.py times -- dropout 0.5 drawn in-kernel -- against the oracle (the reference's dense
    formulation) fed the very masks the kernels draw: loss <= 1e-5, every gradient <= 1e-4."""
    opt = synth.tacred_opt(vocab_size=800, cuda=True, gemm_mode=gemm_mode)
    torch.manual_seed(13)
    tr = GCNTrainer(dict(opt))
    tr.model.train()
    oracle = gcn_oracle.DenseClassifier(dict(opt, cuda=False))
    oracle.load_state_dict({k: v.detach().cpu() for k, v in tr.model.state_dict().items()})
    oracle.train()
    batch = synth.make_batch(33, batch_size=50, vocab_size=800)
    fused = FusedTrainStep(tr)
    dev = [t.to(DEV) for t in batch[:-2]]
    csr = ops.prune_csr(dev[5], dev[6], dev[7], dev[4], dev[1], opt['prune_k'])
    masks = _materialised_masks(tr, batch, csr)
    real = (~batch[1]).unsqueeze(2).expand_as(masks['in'])        # padded tokens gather the all-zero <PAD> row
    keep = float((masks['in'][real] != 0).float().mean())
    assert 0.47 < keep < 0.53
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    loss, _, got = fused.gradients(batch)
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    checked = 0
    for name, p in oracle.named_parameters():
        if p.grad is None:
            continue
        key = name if name in got else name.replace('gcn_model.gcn.', 'gcn_model.')
        assert key in got, name
        assert _rel(got[key], p.grad) <= 1e-4, name
        checked += 1
    assert checked >= 8


def test_two_training_forwards_before_one_backward_use_the_first_forwards_masks():
    """the backward kernels re-derive the Philox masks from {seed, step}; a second training forward advances
    the live counter, so each forward freezes its own copy."""
    over = dict(vocab_size=600, cuda=True, gemm_mode='fp32')
    b1 = synth.make_batch(81, batch_size=40, vocab_size=600)
    b2 = synth.make_batch(82, batch_size=40, vocab_size=600)
    a, b = _pair(over, seed=21)
    b.fast_update = False                       # both on the autograd path
    a.optimizer.zero_grad()
    a.update(b1).backward()
    b.optimizer.zero_grad()
    l1 = b.update(b1)
    b.update(b2)                                # advances the dropout step before l1's backward runs
    l1.backward()
    for (n, pa), (_, pb) in zip(a.model.named_parameters(), b.model.named_parameters()):
        if pa.grad is not None:          # identical masks; float atomics in K5's backward may reorder the last bits
            assert _rel(pb.grad, pa.grad) <= 1e-6, n


@pytest.mark.parametrize('optim', ('sgd', 'adagrad', 'adam', 'adamax'))
def test_train_step_runs_for_every_optimizer_the_reference_cli_offers(optim):
    """Adam / Adamax are built capturable on CUDA, so GraphedTrainStep's capture of optimizer.step() holds."""
    tr = GCNTrainer(synth.tacred_opt(vocab_size=500, cuda=True, optim=optim, lr=0.3 if optim == 'sgd' else 0.01,
                                     input_dropout=0.0, gcn_dropout=0.0))
    tr.model.train()
    batch = synth.make_batch(8, batch_size=50, vocab_size=500)
    losses = [float(tr.train_step(batch)) for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    assert tr._graphed.replays >= 3
    assert isinstance(tr._graphed, FusedTrainStep if optim == 'sgd' else GraphedTrainStep)


def test_graphed_engine_leaves_the_reference_sequence_with_a_dense_embedding_gradient():
    """the row-sparse routing is scoped to GraphedTrainStep's own step."""
    tr = GCNTrainer(synth.tacred_opt(vocab_size=500, cuda=True, conv_l2=1e-4, input_dropout=0.0, gcn_dropout=0.0))
    tr.model.train()
    batch = synth.make_batch(9, batch_size=50, vocab_size=500)
    for _ in range(5):
        tr.train_step(batch)
    assert isinstance(tr._graphed, GraphedTrainStep) and tr._graphed.sparse is not None
    assert tr.model.gcn_model.gcn.sparse_embedding is None
    assert float(tr._graphed.sparse.G.abs().max()) == 0.0
    emb0 = tr.model.gcn_model.emb.weight.detach().clone()
    tr.optimizer.zero_grad()
    loss = tr.update(batch)
    loss.backward()
    g = tr.model.gcn_model.emb.weight.grad
    assert g is not None and float(g.abs().max()) > 0
    torch.nn.utils.clip_grad_norm_(tr.model.parameters(), 5.0)
    tr.optimizer.step()
    assert not torch.equal(tr.model.gcn_model.emb.weight, emb0)
    assert float(tr._graphed.sparse.G.abs().max()) == 0.0


# ---------------------------------------------------------------- K11 / FusedPredict ----------------------------------

@pytest.mark.parametrize('B,C', [(50, 42), (1, 42), (20, 19), (333, 10), (64, 200)])
def test_k11_predict_tail_matches_torch(B, C):
    """softmax / np.argmax (first maximum) / mean CE / un-sort of model/trainer.py:118-123 in one launch."""
    g = torch.Generator().manual_seed(B * 31 + C)
    logits = torch.randn(B, C, generator=g) * 3
    logits[0, :] = 1.25                                  # a full tie: numpy picks column 0
    if B > 2:
        logits[2, 5 % C] = logits[2, C - 1] = 9.0        # a two-way tie: the lower column wins
    labels = torch.randint(0, C, (B,), generator=g)
    orig_idx = torch.randperm(B, generator=g).tolist()
    order = sorted(range(B), key=orig_idx.__getitem__)
    dest = torch.empty(B, dtype=torch.int32)
    dest[order] = torch.arange(B, dtype=torch.int32)
    result = torch.zeros(ops.predict_result_bytes(B, C), dtype=torch.uint8, device=DEV)
    ops.predict_tail(logits.to(DEV), labels.to(DEV), dest.to(DEV), result)
    raw = result.cpu().numpy()
    probs = raw[:B * C * 4].view(np.float32).reshape(B, C)
    preds = raw[B * C * 4:B * C * 4 + B * 4].view(np.int32)
    loss = float(raw[B * C * 4 + B * 4:].view(np.float32)[0])
    ref_probs = torch.softmax(logits, 1).numpy()
    ref_preds = np.argmax(logits.numpy(), axis=1)
    _, ref_preds_u, ref_probs_u = [list(t) for t in zip(*sorted(zip(orig_idx, ref_preds.tolist(), ref_probs.tolist())))]
    assert preds.tolist() == ref_preds_u
    assert np.abs(probs - np.array(ref_probs_u)).max() <= 1e-6
    ref_loss = float(torch.nn.functional.cross_entropy(logits, labels))
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)


@pytest.mark.parametrize('gemm_mode', ('fp32', 'tf32x3'))
def test_fused_predict_equals_the_per_op_predict(gemm_mode):
    """GCNTrainer.predict on one graph replay + one D2H copy == the same call on the per-op path (model forward, ATen
    softmax / argmax, two copies, Python sort): same labels, probabilities <= 1e-5, loss <= 1e-5; before and after the
    training engines have re-homed the parameters; unsort=False; repeated shapes (replays)."""
    from gcn_over_pruned_trees_b200.engine import FusedPredict
    opt = synth.tacred_opt(vocab_size=700, cuda=True, gemm_mode=gemm_mode)
    torch.manual_seed(17)
    tr = GCNTrainer(dict(opt))
    batches = [synth.make_batch(90 + i, batch_size=bs, vocab_size=700) for i, bs in enumerate((50, 50, 37, 1))]
    batches.append(batches[0])

    def both(b, unsort=True):
        tr.fast_update = False
        ref = tr.predict(b, unsort)
        tr.fast_update = True
        got = tr.predict(b, unsort)
        assert got[0] == ref[0]
        assert np.abs(np.array(got[1]) - np.array(ref[1])).max() <= 1e-5 * np.abs(np.array(ref[1])).max()
        assert abs(got[2] - ref[2]) <= 1e-5 * abs(ref[2])
        assert isinstance(got[0][0], int) and isinstance(got[1][0][0], float) and isinstance(got[2], float)

    for b in batches:
        both(b)
    both(batches[0], unsort=False)
    assert isinstance(tr._fused_predict, FusedPredict) and tr._fused_predict.replays >= 1
    tr.model.train()
    for _ in range(3):                                   # FastUpdate / FusedTrainStep move the parameters into one buffer
        _five_calls(tr, batches[0])
    tr.train_step(batches[2])
    for b in batches[:3]:
        both(b)
    assert not tr.model.training                         # predict() leaves the model in eval mode, as the reference does
