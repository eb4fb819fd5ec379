"""CPU: the SOURCE of csrc/prune_csr.cu (K1, batched path-centric pruning -> CSR) executed on the host (tests/emu: one
fiber per CUDA thread, __syncwarp / shuffles / ballot / match_any as warp barriers) through ops.prune_csr and the C ABI
signature of _lib.py, bit-exact against the real reference's adjacency (tests/golden/adjacency.npz) and the oracle.
The `-m gpu` tests of test_gpu_parity.py run the same cases on the device."""
import ctypes
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

import cases
from gcn_over_pruned_trees_b200 import _lib, ops, synth
from oracle import tree_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))


@pytest.fixture(scope='module', autouse=True)
def emulated():
    import emu_build
    handle = ctypes.CDLL(emu_build.build())
    handle.gpt_prune_csr.argtypes = _lib.SIGNATURES['gpt_prune_csr']
    handle.gpt_prune_csr.restype = ctypes.c_int
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, '_lib', handle)
    mp.setattr(ops, '_dev', lambda t, dtype, name: t.contiguous() if t.dtype == dtype else (_ for _ in ()).throw(
        TypeError('%s must be %s' % (name, dtype))))
    mp.setattr(ops, '_stream', lambda: None)
    yield handle
    mp.undo()


def _csr_of(batch, k):
    return ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], k)


@pytest.fixture(autouse=True, params=['warp_per_sentence', 'cta_per_sentence'])
def k1_form(request, monkeypatch):
    """Every case below runs through both forms of K1: one warp per sentence (the default below 256 tokens) and one CTA
    per sentence (the default from 256 tokens up); gpt_prune_csr reads the threshold on every call."""
    monkeypatch.setenv('GPT_K1_BLOCK_MIN_T', '1' if request.param == 'cta_per_sentence' else '1000000')
    return request.param


def _oracle_adj(batch, k):
    lens = synth.batch_lengths(batch).numpy()
    return tree_oracle.batch_adjacency(batch[5].numpy(), batch[6].numpy(), batch[7].numpy(), batch[4].numpy(), lens,
                                       k, batch[0].shape[1])


@pytest.mark.parametrize('split', cases.SPLITS)
def test_k1_source_on_bundled_sample_matches_reference(golden_adj, split):
    batch = cases.batch_from_npz(golden_adj, split)
    for k in cases.PRUNE_KS:
        csr = _csr_of(batch, k)
        assert int(csr.err.abs().sum()) == 0
        want = golden_adj['%s/adj_k%d' % (split, k)].astype(np.float32)
        assert np.array_equal(csr.to_dense().numpy(), want), (split, k)
        nz = want != 0
        assert np.array_equal(csr.denom.numpy(), nz.sum(2) + 1.0)
        assert np.array_equal((csr.flags.numpy() & 1) != 0, (nz.sum(2) + nz.sum(1)) != 0)
        assert np.array_equal((csr.flags.numpy() & 2) != 0, batch[6].numpy() == 0)
        assert np.array_equal((csr.flags.numpy() & 4) != 0, batch[7].numpy() == 0)
        assert np.array_equal(csr.lens.numpy(), synth.batch_lengths(batch).numpy())


@pytest.mark.parametrize('seed', cases.SYNTH_ADJ_SEEDS[:4])
def test_k1_source_on_synthetic_batches_matches_reference_digests(golden_adj, seed):
    batch = synth.make_batch(seed, batch_size=50)
    for k in cases.PRUNE_KS:
        csr = _csr_of(batch, k)
        got = csr.to_dense().numpy()
        sha = hashlib.sha256(got.astype(np.uint8).tobytes()).hexdigest()[:16]
        assert sha == bytes(golden_adj['synth/%d/k%d/sha' % (seed, k)]).decode(), (seed, k)
        rp, col = csr.rowptr.numpy(), csr.col.numpy()
        for b in range(0, 50, 7):                                   # canonical CSR: columns ascending inside a row
            for t in range(batch[0].shape[1]):
                assert np.all(np.diff(col[b, rp[b, t]:rp[b, t + 1]]) > 0)
    assert np.array_equal(_csr_of(batch, 7).to_dense().numpy(), _oracle_adj(batch, 7))


@pytest.mark.parametrize('name', sorted(cases.EDGE_TREES))
def test_k1_source_on_edge_trees(golden_adj, name):
    head, subj, obj, deprel = cases.EDGE_TREES[name]
    n, width = len(head), len(head) + 3           # padded on purpose
    pad = lambda a, fill=0: torch.tensor([list(a) + [fill] * (width - n)], dtype=torch.int64)
    masks = torch.tensor([[False] * n + [True] * (width - n)])
    sp = torch.from_numpy(cases.positions(subj, n, width=width))[None]
    op = torch.from_numpy(cases.positions(obj, n, width=width))[None]
    for k in cases.PRUNE_KS:
        csr = ops.prune_csr(pad(head), sp, op, pad(deprel), masks, k)
        assert (int(csr.err[0]) & ops.TREE_ERR_FATAL) == 0, (name, k, int(csr.err[0]))
        dense = csr.to_dense().numpy()
        assert np.array_equal(dense[0, :n, :n], golden_adj['edge/%s/k%d' % (name, k)].astype(np.float32)), (name, k)
        assert dense[0, n:, :].sum() == 0


def test_k1_source_on_512_token_sentences(golden_adj):
    batch = synth.make_batch(900, batch_size=6, fixed_len=512)
    for k in (-1, 1):
        got = _csr_of(batch, k).to_dense().numpy()
        sha = hashlib.sha256(got.astype(np.uint8).tobytes()).hexdigest()[:16]
        assert sha == bytes(golden_adj['synth512/k%d/sha' % k]).decode()


def test_k1_source_flags_malformed_trees_instead_of_hanging():
    def run(head, subj, obj, k, deprel=None):
        n = len(head)
        t = lambda a: torch.tensor([a], dtype=torch.int64)
        csr = ops.prune_csr(t(head), torch.from_numpy(cases.positions(subj, n))[None],
                            torch.from_numpy(cases.positions(obj, n))[None], t(deprel or [5] * n),
                            torch.zeros((1, n), dtype=torch.bool), k)
        return int(csr.err[0]), csr
    e, csr = run([2, 3, 1], [0], [2], 1)
    assert e & 4 and csr.to_dense().sum() == 0                       # cycle (reference: infinite loop)
    assert run([2, 3, 1], [0], [2], -1)[0] & (2 | 4)                 # no root
    assert run([0, 0], [0], [1], 0)[0] & 16                          # entities under different roots
    assert run([0, 1], [], [1], 0)[0] & 8                            # empty subject span
    assert run([0, 5], [0], [1], 0)[0] & 1                           # head out of range
    assert run([0, 1], [0], [1], 0, deprel=[11, 250])[0] & 32        # deprel does not fit uint8 + 42
    assert run([0, 1, 1, 2], [3], [2], 1, deprel=[11, 0, 5, 0])[0] == 64   # warning only
