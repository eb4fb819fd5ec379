"""Device-resident loader against batches recorded from the UNMODIFIED reference loader
(tests/golden/make_loader_golden.py -> tests/golden/loader.npz).  CPU part: the host-side plan (ids, order, widths);
GPU part: the tensors K9 writes, bit-exact."""
import json
import os
import random

import numpy as np
import pytest
import torch

from gcn_over_pruned_trees_b200 import constant
from gcn_over_pruned_trees_b200.data import loader as dloader
from oracle import loader_oracle
from tests.golden import make_loader_golden as gold_src

HERE = os.path.dirname(os.path.abspath(__file__))
SAMPLE = os.path.join(HERE, 'golden', 'loader_sample.json')
SAMPLE_SEMEVAL = os.path.join(HERE, 'golden', 'loader_sample_semeval.json')
GOLD = np.load(os.path.join(HERE, 'golden', 'loader.npz'))
TACRED_NAMES = ('words', 'masks', 'pos', 'ner', 'deprel', 'head', 'subj_pos', 'obj_pos', 'rels')
SEMEVAL_NAMES = ('words', 'masks', 'pos', 'deprel', 'head', 'subj_pos', 'obj_pos', 'rels')


class Vocab(object):
    def __init__(self, words):
        self.id2word = words
        self.word2id = {w: i for i, w in enumerate(words)}
        self.size = len(words)


VOCAB = Vocab(gold_src.vocab_words())


# ------------------------------------------------------------------------------------------------ CPU: host logic --

def _oracle_check(prefix, sample, opt, evaluation, names, with_ner, label2id, shuffle_seed=None, np_seed=None):
    data = dloader.preprocess(json.load(open(sample)), VOCAB.word2id, opt, label2id, with_ner=with_ner)
    if shuffle_seed is not None:
        random.seed(shuffle_seed)
        idx = list(range(len(data)))
        random.shuffle(idx)
        data = [data[i] for i in idx]
    if np_seed is not None:
        np.random.seed(np_seed)
    n_batches = int(GOLD[prefix + '/n_batches'])
    assert n_batches == (len(data) + 15) // 16
    for k in range(n_batches):
        batch = loader_oracle.get_batch(data[16 * k:16 * k + 16], evaluation, opt['word_dropout'], with_ner=with_ner)
        assert len(batch) == len(names) + 1
        for name, t in zip(names, batch):
            assert np.array_equal(t.numpy().astype(np.int64), GOLD['%s/%d/%s' % (prefix, k, name)].astype(np.int64)), \
                (prefix, k, name)
        assert list(batch[-1]) == GOLD['%s/%d/orig_idx' % (prefix, k)].tolist()


def test_oracle_and_host_preprocessing_reproduce_the_reference_batches():
    """preprocess (ids, entity masking, <UNK> mapping, positions) + the oracle's restatement of __getitem__ (length sort
    with the reference's tie-break, numpy word dropout, padding) against the unmodified reference loader's output."""
    _oracle_check('eval', SAMPLE, dict(lower=False, word_dropout=0.04), True, TACRED_NAMES, True, constant.LABEL_TO_ID)
    _oracle_check('train', SAMPLE, dict(lower=True, word_dropout=0.2), False, TACRED_NAMES, True,
                  constant.LABEL_TO_ID, shuffle_seed=5, np_seed=7)
    _oracle_check('semeval', SAMPLE_SEMEVAL, dict(lower=False, word_dropout=0.0), True, SEMEVAL_NAMES, False,
                  dloader.SEMEVAL_LABEL_TO_ID)


def test_sorted_rows_is_the_reference_tie_break():
    lens = [5, 9, 5, 9, 3]
    assert dloader.sorted_rows(lens) == [3, 1, 2, 0, 4] == loader_oracle.sort_all([lens], lens)[1]


def test_unknown_words_and_tags_map_to_unk_and_entities_are_masked():
    d = json.load(open(SAMPLE))[0]
    row = dloader.preprocess([d], VOCAB.word2id, dict(lower=False), constant.LABEL_TO_ID)[0]
    for t in range(d['subj_start'], d['subj_end'] + 1):
        assert row[0][t] == VOCAB.word2id['SUBJ-' + d['subj_type']] and row[5][t] == 0
    for t in range(d['obj_start'], d['obj_end'] + 1):
        assert row[0][t] == VOCAB.word2id['OBJ-' + d['obj_type']] and row[6][t] == 0
    for t, w in enumerate(d['token']):
        if w in ('OOVword', 'Zyx') and row[5][t] != 0 and row[6][t] != 0:
            assert row[0][t] == constant.UNK_ID
    assert dloader.get_positions(2, 3, 6) == [-2, -1, 0, 0, 1, 2]


def test_loader_without_a_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from gcn_over_pruned_trees_b200._lib import GptError
    with pytest.raises(GptError):
        dloader.DataLoader(SAMPLE, 16, dict(lower=False, word_dropout=0.0), VOCAB, evaluation=True)


# ------------------------------------------------------------------------------------------------ GPU: K9 ---------

def _check(batch, prefix, k, names):
    for name, t in zip(names, batch[:len(names)]):
        assert t.is_cuda
        want = GOLD['%s/%d/%s' % (prefix, k, name)]
        got = t.cpu().numpy()
        assert got.dtype == (np.bool_ if name == 'masks' else np.int64), name
        assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), (prefix, k, name)
    assert list(batch[-1]) == GOLD['%s/%d/orig_idx' % (prefix, k)].tolist()


@pytest.mark.gpu
def test_eval_batches_equal_the_reference_loader_bit_for_bit():
    dl = dloader.DataLoader(SAMPLE, 16, dict(lower=False, word_dropout=0.04), VOCAB, evaluation=True)
    assert len(dl) == int(GOLD['eval/n_batches']) and dl.num_examples == 37
    for k, batch in enumerate(dl):
        assert len(batch) == 10
        _check(batch, 'eval', k, TACRED_NAMES)
    with pytest.raises(IndexError):
        dl[len(dl)]
    with pytest.raises(TypeError):
        dl['0']


@pytest.mark.gpu
def test_train_batches_with_the_reference_numpy_dropout_stream_equal_the_reference():
    """Same random.seed -> same shuffle; host_word_dropout replays np.random exactly as loader.py:181-188 consumes it."""
    random.seed(5)
    np.random.seed(7)
    dl = dloader.DataLoader(SAMPLE, 16, dict(lower=True, word_dropout=0.2), VOCAB, evaluation=False,
                            host_word_dropout=True)
    assert [constant.LABEL_TO_ID[x] for x in dl.gold()] == GOLD['train/gold'].tolist()
    for k in range(len(dl)):
        _check(dl[k], 'train', k, TACRED_NAMES)


@pytest.mark.gpu
def test_semeval_batches_are_9_tuples_equal_to_the_reference():
    dl = dloader.DataLoader(SAMPLE_SEMEVAL, 16, dict(lower=False, word_dropout=0.0, dataset='semeval'), VOCAB,
                            evaluation=True)
    for k, batch in enumerate(dl):
        assert len(batch) == 9
        _check(batch, 'semeval', k, SEMEVAL_NAMES)


@pytest.mark.gpu
def test_device_word_dropout_rate_and_rule():
    """Device Philox dropout: only words change, only into <UNK>, <UNK>/<PAD> never change, rate ~ p, and two passes
    over the data draw different masks."""
    random.seed(1)
    opt = dict(lower=False, word_dropout=0.3)
    dl = dloader.DataLoader(SAMPLE, 37, opt, VOCAB, evaluation=False, seed=123)
    random.seed(1)
    clean = dloader.DataLoader(SAMPLE, 37, dict(opt, word_dropout=0.0), VOCAB, evaluation=False, seed=123)
    ref = clean[0]
    changed, eligible, masks_seen = 0, 0, []
    for rep in range(40):
        b = dl[0]
        for i in (1, 2, 3, 4, 5, 6, 7, 8):
            assert torch.equal(b[i], ref[i])
        w, w0 = b[0], ref[0]
        diff = w != w0
        assert bool((w[diff] == constant.UNK_ID).all())
        assert not bool(diff[(w0 == constant.UNK_ID) | (w0 == constant.PAD_ID)].any())
        changed += int(diff.sum())
        eligible += int(((w0 != constant.UNK_ID) & (w0 != constant.PAD_ID)).sum())
        masks_seen.append(diff.cpu())
    assert abs(changed / eligible - 0.3) < 0.03
    assert not torch.equal(masks_seen[0], masks_seen[1])


@pytest.mark.gpu
def test_loader_batches_drive_the_fused_training_step():
    """loader -> PackedBatch on the device -> FusedTrainStep: the path bench.py's e2e.device_loader times."""
    from gcn_over_pruned_trees_b200 import synth
    from gcn_over_pruned_trees_b200.engine import FusedTrainStep
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
    random.seed(3)
    dl = dloader.DataLoader(SAMPLE, 16, dict(lower=False, word_dropout=0.04), VOCAB, evaluation=False, seed=9)
    torch.manual_seed(0)
    tr = GCNTrainer(synth.tacred_opt(vocab_size=VOCAB.size, cuda=True))
    tr.model.train()
    eng = FusedTrainStep(tr)
    losses = [float(eng.step_from(dl, k % len(dl))) for k in range(15)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert eng.replays >= 3                              # the later steps ran from K9 straight into the graph's buffers
    a = dl.packed(0)
    b = dl[0]
    assert all(torch.equal(x, y) for x, y in zip(a.as_tuple()[1:9], b[1:9]))     # ring buffer vs owned tuple
