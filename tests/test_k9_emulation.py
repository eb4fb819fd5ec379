"""CPU: the SOURCE of csrc/batch.cu (K9, device-resident batch builder) executed on the host (tests/emu) underneath
data/loader.py's DataLoader, bit-exact against batches recorded from the UNMODIFIED reference loader
(tests/golden/loader.npz).  The `-m gpu` tests of test_loader.py run the same cases on the device."""
import ctypes
import os
import random
import sys

import numpy as np
import pytest
import torch

from gcn_over_pruned_trees_b200 import _lib, constant, ops
from gcn_over_pruned_trees_b200.data import loader as dloader
from tests.test_loader import GOLD, SAMPLE, SAMPLE_SEMEVAL, SEMEVAL_NAMES, TACRED_NAMES, VOCAB

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))


@pytest.fixture(scope='module', autouse=True)
def emulated():
    import emu_build
    handle = ctypes.CDLL(emu_build.build())
    handle.gpt_build_batch.argtypes = _lib.SIGNATURES['gpt_build_batch']
    handle.gpt_build_batch.restype = ctypes.c_int
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, '_lib', handle)
    mp.setattr(ops, '_stream', lambda: None)
    yield handle
    mp.undo()


def _check(batch, prefix, k, names):
    for name, t in zip(names, batch[:len(names)]):
        want = GOLD['%s/%d/%s' % (prefix, k, name)]
        got = t.numpy()
        assert got.dtype == (np.bool_ if name == 'masks' else np.int64), name
        assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), (prefix, k, name)
    assert list(batch[-1]) == GOLD['%s/%d/orig_idx' % (prefix, k)].tolist()


def test_k9_source_eval_batches_equal_the_reference_loader_bit_for_bit():
    dl = dloader.DataLoader(SAMPLE, 16, dict(lower=False, word_dropout=0.04), VOCAB, evaluation=True, device='cpu')
    assert len(dl) == int(GOLD['eval/n_batches']) and dl.num_examples == 37
    for k, batch in enumerate(dl):
        assert len(batch) == 10
        _check(batch, 'eval', k, TACRED_NAMES)


def test_k9_source_train_batches_with_the_reference_numpy_dropout_stream():
    random.seed(5)
    np.random.seed(7)
    dl = dloader.DataLoader(SAMPLE, 16, dict(lower=True, word_dropout=0.2), VOCAB, evaluation=False,
                            host_word_dropout=True, device='cpu')
    assert [constant.LABEL_TO_ID[x] for x in dl.gold()] == GOLD['train/gold'].tolist()
    for k in range(len(dl)):
        _check(dl[k], 'train', k, TACRED_NAMES)


def test_k9_source_semeval_batches_are_9_tuples_equal_to_the_reference():
    dl = dloader.DataLoader(SAMPLE_SEMEVAL, 16, dict(lower=False, word_dropout=0.0, dataset='semeval'), VOCAB,
                            evaluation=True, device='cpu')
    for k, batch in enumerate(dl):
        assert len(batch) == 9
        _check(batch, 'semeval', k, SEMEVAL_NAMES)


def test_k9_source_word_dropout_rate_and_rule():
    random.seed(1)
    opt = dict(lower=False, word_dropout=0.3)
    dl = dloader.DataLoader(SAMPLE, 37, opt, VOCAB, evaluation=False, seed=123, device='cpu')
    random.seed(1)
    clean = dloader.DataLoader(SAMPLE, 37, dict(opt, word_dropout=0.0), VOCAB, evaluation=False, seed=123,
                               device='cpu')
    ref = clean[0]
    changed, eligible, masks_seen = 0, 0, []
    for rep in range(20):
        b = dl[0]
        for i in (1, 2, 3, 4, 5, 6, 7, 8):
            assert torch.equal(b[i], ref[i])
        w, w0 = b[0], ref[0]
        diff = w != w0
        assert bool((w[diff] == constant.UNK_ID).all())
        assert not bool(diff[(w0 == constant.UNK_ID) | (w0 == constant.PAD_ID)].any())
        changed += int(diff.sum())
        eligible += int(((w0 != constant.UNK_ID) & (w0 != constant.PAD_ID)).sum())
        masks_seen.append(diff)
    assert abs(changed / eligible - 0.3) < 0.04
    assert not torch.equal(masks_seen[0], masks_seen[1])
