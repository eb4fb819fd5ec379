"""CPU: pin the oracle (oracle/*.py) to outputs of the real reference stored in tests/golden/*.npz."""
import hashlib

import numpy as np
import pytest
import torch

import cases
import weights
from gcn_over_pruned_trees_b200 import synth
from oracle import gcn_oracle, tree_oracle


def _sha(adj):
    return hashlib.sha256(adj.astype(np.uint8).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize('split', cases.SPLITS)
@pytest.mark.parametrize('k', cases.PRUNE_KS)
def test_bundled_sample_adjacency_bit_exact(golden_adj, split, k):
    b = cases.batch_from_npz(golden_adj, split)
    lens = synth.batch_lengths(b).numpy()
    got = tree_oracle.batch_adjacency(b[5].numpy(), b[6].numpy(), b[7].numpy(), b[4].numpy(), lens, k)
    want = golden_adj['%s/adj_k%d' % (split, k)]
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want.astype(np.float32))


def test_survey_known_answer_sentence0():
    # sentence 0 of train.json, k=0 -> LCA 14, kept {12,14,16,20}, 10 COO entries
    head = [2, 0, 4, 2, 6, 4, 11, 9, 11, 11, 6, 15, 15, 15, 11, 17, 15, 21, 21, 21, 17, 2]
    deprel = [7, 11, 14, 21, 14, 23, 14, 13, 7, 14, 21, 14, 7, 20, 21, 6, 10, 14, 24, 4, 25, 2]
    n = len(head)
    adj = tree_oracle.pruned_adjacency(np.array(head), cases.positions([20], n), cases.positions([12], n),
                                       np.array(deprel), n, 0, n)
    coo = sorted((int(i), int(j), int(adj[i, j])) for i, j in zip(*np.nonzero(adj)))
    assert coo == [(12, 12, 84), (12, 14, 49), (14, 12, 7), (14, 14, 84), (14, 16, 10), (16, 14, 52),
                   (16, 16, 84), (16, 20, 25), (20, 16, 67), (20, 20, 84)]
    for k, nnz in ((-1, 64), (1, 28), (2, 28)):
        a = tree_oracle.pruned_adjacency(np.array(head), cases.positions([20], n), cases.positions([12], n),
                                         np.array(deprel), n, k, n)
        assert int((a != 0).sum()) == nnz


@pytest.mark.parametrize('seed', cases.SYNTH_ADJ_SEEDS)
def test_synthetic_adjacency_digest(golden_adj, seed):
    b = synth.make_batch(seed, batch_size=50)
    lens = synth.batch_lengths(b).numpy()
    for k in cases.PRUNE_KS:
        got = tree_oracle.batch_adjacency(b[5].numpy(), b[6].numpy(), b[7].numpy(), b[4].numpy(), lens, k)
        assert _sha(got) == bytes(golden_adj['synth/%d/k%d/sha' % (seed, k)]).decode()
        assert int(got.sum()) == int(golden_adj['synth/%d/k%d/sum' % (seed, k)])


def test_synthetic_512_token_digest(golden_adj):
    b = synth.make_batch(900, batch_size=6, fixed_len=512)
    for k in (-1, 1):
        got = tree_oracle.batch_adjacency(b[5].numpy(), b[6].numpy(), b[7].numpy(), b[4].numpy(), [512] * 6, k)
        assert _sha(got) == bytes(golden_adj['synth512/k%d/sha' % k]).decode()


@pytest.mark.parametrize('name', sorted(cases.EDGE_TREES))
def test_edge_case_trees(golden_adj, name):
    head, subj, obj, deprel = cases.EDGE_TREES[name]
    n = len(head)
    for k in cases.PRUNE_KS:
        got = tree_oracle.pruned_adjacency(np.array(head), cases.positions(subj, n), cases.positions(obj, n),
                                           np.array(deprel), n, k, n)
        assert np.array_equal(got, golden_adj['edge/%s/k%d' % (name, k)].astype(np.float32)), (name, k)


def test_structural_invariants():
    # nnz == 3*n_kept - 2, diag == 84 <=> kept (with >= 1 neighbour), symmetric support
    b = synth.make_batch(5, batch_size=40)
    lens = synth.batch_lengths(b).numpy()
    for k in (-1, 0, 1, 2):
        adj = tree_oracle.batch_adjacency(b[5].numpy(), b[6].numpy(), b[7].numpy(), b[4].numpy(), lens, k)
        for a in adj:
            kept = int((np.diag(a) == 84).sum())
            assert int((a != 0).sum()) == 3 * kept - 2
            assert np.array_equal(a != 0, (a != 0).T)
            assert a.max() <= 84


def test_malformed_inputs_raise():
    with pytest.raises(tree_oracle.MalformedTree):      # cycle: reference never returns
        tree_oracle.pruned_adjacency(np.array([2, 1, 0]), cases.positions([0], 3), cases.positions([2], 3),
                                     np.array([2, 3, 11]), 3, 1, 3)
    with pytest.raises(tree_oracle.MalformedTree):      # entities under different roots
        tree_oracle.pruned_adjacency(np.array([0, 0]), cases.positions([0], 2), cases.positions([1], 2),
                                     np.array([11, 11]), 2, 0, 2)
    with pytest.raises(tree_oracle.MalformedTree):      # empty subject span
        tree_oracle.pruned_adjacency(np.array([0, 1]), cases.positions([], 2), cases.positions([1], 2),
                                     np.array([11, 2]), 2, 0, 2)


def _case_setup(golden_adj, name, table=None):
    over, source, wseed = (cases.MODEL_CASES if table is None else table)[name]
    if source[0] == 'split':
        batch = cases.batch_from_npz(golden_adj, source[1])
        over = dict(over, vocab_size=int(golden_adj['vocab_size']))
    else:
        batch = cases.make_case_batch(source, over)
    opt = synth.tacred_opt(**over)
    model = gcn_oracle.DenseClassifier(opt)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()})
    return opt, batch, model


@pytest.mark.parametrize('name', sorted(cases.MODEL_CASES))
def test_dense_model_matches_reference_eval(golden_adj, golden_model, name):
    opt, batch, model = _case_setup(golden_adj, name)
    model.eval()
    with torch.no_grad():
        loss, logits = model.loss(batch)
    want = golden_model['%s/logits' % name]
    assert np.abs(logits.numpy() - want).max() <= 1e-6 * np.abs(want).max()
    assert abs(loss.item() - float(golden_model['%s/eval_loss' % name])) <= 1e-6 * abs(loss.item())


@pytest.mark.parametrize('name', cases.GRAD_CASES)
def test_dense_model_matches_reference_train_grads(golden_adj, golden_model, name):
    opt, batch, model = _case_setup(golden_adj, name)
    model.train()
    torch.manual_seed(cases.DROPOUT_SEED)     # same draw order as the reference: in_drop, (rnn_drop,) gcn_drop
    loss, _ = model.loss(batch)
    loss.backward()
    assert abs(loss.item() - float(golden_model['%s/train_loss' % name])) <= 1e-6 * abs(loss.item())
    seen = set()
    checked = 0
    for key, p in model.named_parameters():
        if p.grad is None or id(p) in seen:
            continue
        seen.add(id(p))
        sample, norm, total = weights.grad_digest(p.grad.numpy())
        want = golden_model['%s/grad/%s/sample' % (name, key)]
        scale = max(float(np.abs(want).max()), 1e-12)
        assert np.abs(sample - want).max() <= 2e-5 * scale, key
        assert abs(norm - float(golden_model['%s/grad/%s/norm' % (name, key)])) <= 1e-5 * max(norm, 1e-12), key
        checked += 1
    assert checked >= 8


# ---- relation-aware adjacency modes: oracle vs tests/golden/deprel.npz (the real reference) ----

_DEPREL_ALL = dict(cases.DEPREL_CASES, **cases.DEPREL_RANDOM_CASES)


@pytest.mark.parametrize('name', sorted(_DEPREL_ALL))
def test_relation_modes_match_reference_eval(golden_adj, golden_deprel, name):
    opt, batch, model = _case_setup(golden_adj, name, _DEPREL_ALL)
    model.eval()
    with torch.no_grad():
        loss, logits = model.loss(batch)
        _, h_out = model(list(batch[:-2]))
    want = golden_deprel['%s/logits' % name]
    assert np.abs(logits.numpy() - want).max() <= 1e-6 * np.abs(want).max()
    want = golden_deprel['%s/h_out' % name]
    assert np.abs(h_out.numpy() - want).max() <= 2e-6 * np.abs(want).max()
    assert abs(loss.item() - float(golden_deprel['%s/eval_loss' % name])) <= 1e-6 * abs(loss.item())


@pytest.mark.parametrize('name', cases.DEPREL_GRAD_CASES + tuple(sorted(cases.DEPREL_RANDOM_CASES)))
def test_relation_modes_match_reference_train_grads(golden_adj, golden_deprel, name):
    """Train mode; every random draw (dropouts, edge dropout, relation forgetting) is taken in the reference's order
    from the same seed, so losses and gradients are comparable."""
    opt, batch, model = _case_setup(golden_adj, name, _DEPREL_ALL)
    model.train()
    torch.manual_seed(cases.DROPOUT_SEED)
    loss, _ = model.loss(batch)
    loss.backward()
    assert abs(loss.item() - float(golden_deprel['%s/train_loss' % name])) <= 1e-6 * abs(loss.item())
    seen, checked = set(), 0
    for key, p in model.named_parameters():
        if p.grad is None or id(p) in seen:
            continue
        seen.add(id(p))
        sample, norm, total = weights.grad_digest(p.grad.numpy())
        want = golden_deprel['%s/grad/%s/sample' % (name, key)]
        scale = max(float(np.abs(want).max()), 1e-12)
        assert np.abs(sample - want).max() <= 2e-5 * scale, key
        assert abs(norm - float(golden_deprel['%s/grad/%s/norm' % (name, key)])) <= 1e-5 * max(norm, 1e-12), key
        checked += 1
    assert checked >= 8
    assert any('deprel_emb' in k for k, p in model.named_parameters() if p.grad is not None)
