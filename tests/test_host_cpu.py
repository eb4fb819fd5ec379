"""CPU tests (-m "not gpu"): the C-ABI library loads and exports what include/gpt_b200.h declares, host-side logic
(generator, batch unpacking, checkpoint layout), and that nothing silently falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import cases
import weights
from gcn_over_pruned_trees_b200 import _lib, constant, ops, synth
from gcn_over_pruned_trees_b200.model import trainer as trainer_mod

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'gpt_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gpt_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run `python __graft_entry__.py build` first'
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(handle, name), name
    assert set(_lib.SIGNATURES) | {'gpt_error_string'} == set(names)
    assert _lib.lib().gpt_version() >= 100
    assert b'bad argument' in _lib.lib().gpt_error_string(-1)


def test_ops_refuse_cpu_tensors():
    b = synth.make_batch(0, batch_size=3)
    with pytest.raises(_lib.GptError):
        ops.prune_csr(b[5], b[6], b[7], b[4], b[1], 1)
    trainer = trainer_mod.GCNTrainer(synth.tacred_opt(vocab_size=50, cuda=False))
    with pytest.raises(_lib.GptError):
        trainer.update(b)


def test_constants_match_reference_id_space():
    assert constant.DEPREL_TO_ID['nsubj'] == 7 and constant.DEPREL_TO_ID['nsubj_reverse'] == 49
    assert constant.DEPREL_TO_ID['self_loop'] == 84 and constant.POS_TO_ID['#'] == 46
    assert constant.NER_TO_ID['SET'] == 14 and constant.LABEL_TO_ID['per:country_of_death'] == 41


def test_synthetic_batch_layout():
    b = synth.make_batch(5, batch_size=50)
    words, masks, pos, ner, deprel, head, subj_pos, obj_pos, rels, orig_idx = b
    lens = synth.batch_lengths(b)
    assert words.shape == (50, int(lens.max())) and torch.all(lens[:-1] >= lens[1:])
    assert 8 <= int(lens.min()) and int(lens.max()) <= 96
    for i in range(50):
        n = int(lens[i])
        assert int((head[i, :n] == 0).sum()) == 1 and torch.all(head[i, n:] == 0)
        assert torch.all(subj_pos[i, n:] == 150) and int((subj_pos[i, :n] == 0).sum()) in (1, 2, 3)
        assert not torch.any((subj_pos[i] == 0) & (obj_pos[i] == 0))
        assert torch.all(deprel[i, :n] >= 2) and torch.all(masks[i, n:])
    s = synth.make_batch(5, batch_size=8, dataset='semeval', num_class=19)
    assert len(s) == 9
    big = synth.make_batch(1, batch_size=3, fixed_len=512)
    assert big[0].shape == (3, 512)


def test_unpack_batch_tacred_and_semeval():
    for ds in ('tacred', 'semeval'):
        b = synth.make_batch(2, batch_size=6, dataset=ds)
        inputs, labels, tokens, head, subj_pos, obj_pos, lens = trainer_mod.unpack_batch(b, False)
        assert len(inputs) == len(b) - 2 and labels is b[-2]
        off = 5 if ds == 'tacred' else 4
        assert head is b[off] and subj_pos is b[off + 1] and obj_pos is b[off + 2]
        assert torch.equal(lens, synth.batch_lengths(b))


@pytest.mark.parametrize('name', ['cfg1_train_json_k1', 'cfg3_cgcn_k1', 'cfg4_semeval_k1', 'sum_pool_3layer_mlp1',
                                  'full_k1_d8', 'full_cgcn_h64', 'full_semeval', 'diag_k1', 'diag_cgcn'])
def test_state_dict_layout_matches_reference_checkpoints(name):
    # the shape tables in weights.py are themselves pinned: make_golden.py / make_deprel_golden.py load states built
    # from them into the REAL reference model with strict=True
    over, _, wseed = dict(cases.MODEL_CASES, **cases.DEPREL_CASES)[name]
    opt = synth.tacred_opt(**dict(over, vocab_size=over.get('vocab_size', 963)))
    model = trainer_mod.GCNTrainer(opt).model
    want = weights.state_shapes(opt)
    for dup in ('emb', 'pos_emb', 'ner_emb', 'deprel_emb'):
        if 'gcn_model.%s.weight' % dup in want:
            want['gcn_model.gcn.%s.weight' % dup] = want['gcn_model.%s.weight' % dup]
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()})  # strict
    sd = model.state_dict()
    assert sd['gcn_model.emb.weight'].data_ptr() == sd['gcn_model.gcn.emb.weight'].data_ptr()


def test_unsupported_paths_fail_loudly():
    with pytest.raises(NotImplementedError):        # builds weights in the reference, but its forward raises (gcn.py:388)
        trainer_mod.GCNTrainer(synth.tacred_opt(vocab_size=50, adj_type='concat_deprel'))
    with pytest.raises(ValueError):                 # the reference dies inside einsum at layer 2
        trainer_mod.GCNTrainer(synth.tacred_opt(vocab_size=50, adj_type='full_deprel', deprel_emb_dim=8))
    with pytest.raises(NotImplementedError):
        trainer_mod.GCNTrainer(synth.tacred_opt(vocab_size=50, emb_dropout=0.1))


def test_relation_modes_have_no_cpu_path():
    from gcn_over_pruned_trees_b200._lib import GptError
    over, source, _ = cases.DEPREL_CASES['full_k1_d8']
    trainer = trainer_mod.GCNTrainer(synth.tacred_opt(**over))
    batch = synth.make_batch(source[1], batch_size=4, vocab_size=over['vocab_size'])
    with pytest.raises(GptError):
        trainer.update(batch)


def test_packed_batch_is_one_buffer_with_the_loader_views():
    """PackedBatch (engine.py): a loader tuple in ONE contiguous buffer -- same shapes, dtypes and values, every view
    naturally aligned, and `like=` reproduces the layout for the static device copy."""
    from gcn_over_pruned_trees_b200.engine import PackedBatch
    for dataset in ('tacred', 'semeval'):
        batch = synth.make_batch(5, batch_size=7, vocab_size=300, dataset=dataset)
        pb = PackedBatch(batch)
        fields = batch[:-2]
        assert len(pb.fields) == len(fields) and pb.key == (7, fields[0].shape[1], len(fields))
        for v, f in zip(pb.fields, fields):
            assert v.dtype == f.dtype and v.shape == f.shape and torch.equal(v, f)
            assert v.data_ptr() % v.element_size() == 0
            lo, hi = pb.buf.data_ptr(), pb.buf.data_ptr() + pb.buf.numel()
            assert lo <= v.data_ptr() and v.data_ptr() + v.numel() * v.element_size() <= hi
        assert torch.equal(pb.labels, batch[-2]) and pb.orig_idx == batch[-1]
        twin = PackedBatch(like=pb)
        twin.buf.copy_(pb.buf)
        assert all(torch.equal(a, b) for a, b in zip(twin.as_tuple()[:-1], pb.as_tuple()[:-1]))


def test_loader_row_order_and_widths_are_planned_on_the_host():
    """data/loader.py: the length sort (reference tie-break) and the batch widths need only the sentence lengths."""
    from gcn_over_pruned_trees_b200.data import loader as dloader
    lens = [3, 7, 7, 2, 9, 7]
    rows = dloader.sorted_rows(lens)
    assert rows == [4, 5, 2, 1, 0, 3]                      # descending length, ties by descending position
    assert [lens[r] for r in rows] == sorted(lens, reverse=True)
    assert dloader.get_positions(0, 0, 4) == [0, 1, 2, 3] and dloader.get_positions(3, 3, 4) == [-3, -2, -1, 0]
